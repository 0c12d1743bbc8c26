"""Count the field operations / 32x32 limb products one tuple costs in each kernel, by running the
instrumented HOST build of the very same per-tuple code (tests/host_arith.cpp).  Writes
profiles/op_counts.json, which bench.py's roofline uses as "work per tuple"."""
import json, os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import numpy as np
import hostlib as H, schnorr_oracle as o, vectors as V

lib = H.build()
tab = H.comb_tables(lib)
rnd = random.Random(3)
sk, nonce, m = rnd.randrange(o.R), rnd.randrange(o.R), rnd.randrange(o.Q)
u, Rp, c = o.sign(sk, nonce, m, mul=V.mul)
pk = V.mul(o.G, sk)
cnt = np.zeros(5, np.uint64)
buf = np.zeros(8, np.uint32)

def measure(fn):
    lib.h_counts_reset(); fn(); lib.h_counts_get(H.ptr(cnt))
    w, mul, sqr, addsub, frm = (int(x) for x in cnt)
    return {"imad_wide_per_tuple": w, "fq_mul": mul, "fq_sqr": sqr, "fq_addsub": addsub, "fr_mont_mul": frm}

out = {}
z1, z2 = rnd.randrange(1, o.Q), rnd.randrange(1, o.Q)
out["verify_affine"] = measure(lambda: lib.h_verify(H.ptr(H.pt_mont(pk)), H.ptr(H.limbs(u)), H.ptr(H.pt_mont(Rp)), H.ptr(H.mont(m)), 1, H.ptr(tab[0]), H.ptr(buf)))
out["verify_projective"] = measure(lambda: lib.h_verify(H.ptr(H.pt_mont(pk, z1)), H.ptr(H.limbs(u)), H.ptr(H.pt_mont(Rp, z2)), H.ptr(H.mont(m)), 0, H.ptr(tab[0]), H.ptr(buf)))
uv = np.zeros(16, np.uint32); uo = np.zeros(8, np.uint32)
out["sign"] = measure(lambda: lib.h_sign(H.ptr(H.limbs(sk)), H.ptr(H.limbs(nonce)), H.ptr(H.mont(m)), H.ptr(tab[0]), H.ptr(uo), H.ptr(uv), H.ptr(buf)))
uvp = np.zeros(16, np.uint32)
out["sign_double"] = measure(lambda: lib.h_sign_double(H.ptr(H.limbs(sk)), H.ptr(H.limbs(nonce)), H.ptr(H.mont(m)), H.ptr(tab[0]), H.ptr(tab[1]), H.ptr(uo), H.ptr(uv), H.ptr(uvp), H.ptr(buf)))
ud, Rd, Rdp, cd = o.sign_double(sk, nonce, m, mul=V.mul)
pkp = V.mul(o.G_NUMS, sk)
out["verify_double_affine"] = measure(lambda: lib.h_verify_double(H.ptr(H.pt_mont(pk)), H.ptr(H.pt_mont(pkp)), H.ptr(H.limbs(ud)), H.ptr(H.pt_mont(Rd)), H.ptr(H.pt_mont(Rdp)), H.ptr(H.mont(m)), 1, H.ptr(tab[0]), H.ptr(tab[1]), H.ptr(buf)))
gen = V.mul(o.G, rnd.randrange(o.R)); pkv = V.mul(gen, sk)
uvg, Rvg, cvg = o.sign_vargen(sk, gen, nonce, m, mul=V.mul)
out["verify_vargen_affine"] = measure(lambda: lib.h_verify_vargen(H.ptr(H.pt_mont(pkv)), H.ptr(H.pt_mont(gen)), H.ptr(H.limbs(uvg)), H.ptr(H.pt_mont(Rvg)), H.ptr(H.mont(m)), 1, H.ptr(buf)))
st = np.concatenate([H.mont(x) for x in range(5)])
out["hades_perm_sparse"] = measure(lambda: lib.h_hades(H.ptr(st), 0))
out["hades_perm_dense"] = measure(lambda: lib.h_hades(H.ptr(st), 1))
out["_note"] = "counted by the instrumented host build of schnorr_b200/csrc (same per-tuple code as the kernels); 1 fq_mul = 112 IMAD.WIDE, 1 fr_mont_mul = 128"
if "--check" not in sys.argv:
    path = os.path.join(ROOT, "profiles", "op_counts.json")
    old = json.load(open(path)) if os.path.exists(path) else {}
    for k, v in out.items():
        if isinstance(v, dict) and isinstance(old.get(k), dict) and "dram_bytes_per_launch_ncu" in old[k]:
            v["dram_bytes_per_launch_ncu"] = old[k]["dram_bytes_per_launch_ncu"]
    json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
