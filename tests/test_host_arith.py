"""The CUDA headers compiled for the CPU (carry-chain blocks emulated): every per-tuple routine the kernels
run, against the oracle.  Catches logic errors without a GPU; the GPU parity tests then only have to
catch PTX-level ones."""
import ctypes
import random

import numpy as np
import pytest

import hostlib as H
import schnorr_oracle as o
import vectors as V

Q, R = o.Q, o.R


@pytest.fixture(scope="module")
def lib():
    return H.build()


@pytest.fixture(scope="module")
def tabs(lib):
    return H.comb_tables(lib)


def test_fq_fr(lib):
    rnd = random.Random(1)
    out = np.zeros(8, np.uint32)
    vals = [0, 1, 2, Q - 1, Q - 2, (1 << 255) % Q, (1 << 32) - 1, 1 << 32, Q >> 1, 1 << 224, (1 << 224) - 1] + \
           [rnd.randrange(Q) for _ in range(60)]
    rinv = pow(H.RADIX, -1, Q)
    for a in vals:
        for b in vals[:24]:
            lib.h_fq_mul(H.ptr(H.limbs(a)), H.ptr(H.limbs(b)), H.ptr(out)); assert H.to_int(out) == a * b * rinv % Q
            lib.h_fq_add(H.ptr(H.limbs(a)), H.ptr(H.limbs(b)), H.ptr(out)); assert H.to_int(out) == (a + b) % Q
            lib.h_fq_sub(H.ptr(H.limbs(a)), H.ptr(H.limbs(b)), H.ptr(out)); assert H.to_int(out) == (a - b) % Q
    for a in vals[1:12]:
        lib.h_fq_inv(H.ptr(H.mont(a)), H.ptr(out)); assert H.unmont(out) == pow(a, -1, Q)
    rv = [0, 1, R - 1, (1 << 250) - 1] + [rnd.randrange(R) for _ in range(30)]
    for a in rv:
        for b in rv[:8]:
            lib.h_fr_mul(H.ptr(H.limbs(a)), H.ptr(H.limbs(b)), H.ptr(out)); assert H.to_int(out) == a * b % R


def test_lazy_dot5_extremes(lib):
    """sum of 5 products with one reduction: worst case 5 (q-1)^2 needs the 17th limb and both final subtractions"""
    rnd = random.Random(7)
    out = np.zeros(8, np.uint32)
    rinv = pow(H.RADIX, -1, Q)
    cases = [([Q - 1] * 5, [Q - 1] * 5), ([Q - 1] * 5, [Q - 2, 1, 0, Q - 1, 2]), ([0] * 5, [Q - 1] * 5), ([1, 0, 0, 0, 0], [5, 0, 0, 0, 0])]
    cases += [([rnd.randrange(Q) for _ in range(5)], [rnd.randrange(Q) for _ in range(5)]) for _ in range(200)]
    cases += [([Q - 1 - rnd.randrange(1 << 40) for _ in range(5)], [Q - 1 - rnd.randrange(1 << 40) for _ in range(5)]) for _ in range(50)]
    for c, s in cases:
        cb = np.concatenate([H.limbs(x) for x in c]); sb = np.concatenate([H.limbs(x) for x in s])
        lib.h_fq_dot5(H.ptr(cb), H.ptr(sb), H.ptr(out))
        assert H.to_int(out) == sum(a * b for a, b in zip(c, s)) * rinv % Q


@pytest.mark.parametrize("dense", [1, 0])
def test_hades(lib, dense):
    rnd = random.Random(2)
    for st in ([0] * 5, [Q - 1] * 5, [rnd.randrange(Q) for _ in range(5)]):
        buf = np.concatenate([H.mont(x) for x in st])
        lib.h_hades(H.ptr(buf), dense)
        assert [H.unmont(buf[8 * i:8 * i + 8]) for i in range(5)] == o.hades_perm(st)


def test_scalar_mul(lib, tabs):
    rnd = random.Random(3)
    uv = np.zeros(16, np.uint32)
    for k in [0, 1, 127, 128, 129, 256, R - 1, R, (1 << 252) - 1, rnd.randrange(R)]:
        lib.h_fixed_mul(H.ptr(tabs[0]), H.ptr(H.limbs(k)), H.ptr(uv))
        assert (H.unmont(uv[:8]), H.unmont(uv[8:])) == V.mul(o.G, k)
        lib.h_fixed_mul(H.ptr(tabs[1]), H.ptr(H.limbs(k)), H.ptr(uv))
        assert (H.unmont(uv[:8]), H.unmont(uv[8:])) == V.mul(o.G_NUMS, k)
    for P in V.torsion_points()[:4] + [o.G, V.rand_curve_point(rnd)]:
        for k in [0, 1, 8, R - 1, R + 5, (1 << 252) - 1, rnd.randrange(1 << 252)]:
            for aff in (1, 0):
                pin = H.pt_mont(P) if aff else H.pt_mont(P, rnd.randrange(1, Q))
                lib.h_var_mul(H.ptr(pin), aff, H.ptr(H.limbs(k)), H.ptr(uv))
                assert (H.unmont(uv[:8]), H.unmont(uv[8:])) == V.mul(P, k)


def test_sign_verify_all_variants(lib, tabs):
    rnd = random.Random(4)
    c, u, uv, uvp = np.zeros(8, np.uint32), np.zeros(8, np.uint32), np.zeros(16, np.uint32), np.zeros(16, np.uint32)
    G, Gp = H.ptr(tabs[0]), H.ptr(tabs[1])
    for _ in range(2):
        sk, nonce, m = rnd.randrange(R), rnd.randrange(R), rnd.randrange(Q)
        z1, z2 = rnd.randrange(1, Q), rnd.randrange(1, Q)
        eu, eR, ec = o.sign(sk, nonce, m, mul=V.mul)
        lib.h_sign(H.ptr(H.limbs(sk)), H.ptr(H.limbs(nonce)), H.ptr(H.mont(m)), G, H.ptr(u), H.ptr(uv), H.ptr(c))
        assert (H.to_int(u), (H.unmont(uv[:8]), H.unmont(uv[8:])), H.to_int(c)) == (eu, eR, ec)
        pk = V.mul(o.G, sk)
        assert lib.h_verify(H.ptr(H.pt_mont(pk)), H.ptr(H.limbs(eu)), H.ptr(H.pt_mont(eR)), H.ptr(H.mont(m)), 1, G, H.ptr(c)) == 1
        assert H.to_int(c) == ec
        assert lib.h_verify(H.ptr(H.pt_mont(pk, z1)), H.ptr(H.limbs(eu)), H.ptr(H.pt_mont(eR, z2)), H.ptr(H.mont(m)), 0, G, H.ptr(c)) == 1
        assert lib.h_verify(H.ptr(H.pt_mont(pk)), H.ptr(H.limbs((eu + 1) % R)), H.ptr(H.pt_mont(eR)), H.ptr(H.mont(m)), 1, G, H.ptr(c)) == 0
        assert lib.h_verify(H.ptr(H.pt_mont(pk)), H.ptr(H.limbs(eu + R)), H.ptr(H.pt_mont(eR)), H.ptr(H.mont(m)), 1, G, H.ptr(c)) == 0
        eu, eR, eRp, ec = o.sign_double(sk, nonce, m, mul=V.mul)
        lib.h_sign_double(H.ptr(H.limbs(sk)), H.ptr(H.limbs(nonce)), H.ptr(H.mont(m)), G, Gp, H.ptr(u), H.ptr(uv), H.ptr(uvp), H.ptr(c))
        assert (H.to_int(u), (H.unmont(uvp[:8]), H.unmont(uvp[8:])), H.to_int(c)) == (eu, eRp, ec)
        pkp = V.mul(o.G_NUMS, sk)
        args = lambda a, b, aff: (H.ptr(H.pt_mont(a, None if aff else z1)), H.ptr(H.pt_mont(b, None if aff else z2)), H.ptr(H.limbs(eu)),
                                  H.ptr(H.pt_mont(eR, None if aff else z2)), H.ptr(H.pt_mont(eRp, None if aff else z1)), H.ptr(H.mont(m)), aff, G, Gp, H.ptr(c))
        assert lib.h_verify_double(*args(pk, pkp, 1)) == 1 and lib.h_verify_double(*args(pk, pkp, 0)) == 1
        assert lib.h_verify_double(*args(pk, pk, 1)) == 0
        gen = V.mul(o.G, rnd.randrange(R))
        eu, eR, ec = o.sign_vargen(sk, gen, nonce, m, mul=V.mul)
        lib.h_sign_vargen(H.ptr(H.limbs(sk)), H.ptr(H.pt_mont(gen)), 1, H.ptr(H.limbs(nonce)), H.ptr(H.mont(m)), H.ptr(u), H.ptr(uv), H.ptr(c))
        assert (H.to_int(u), (H.unmont(uv[:8]), H.unmont(uv[8:])), H.to_int(c)) == (eu, eR, ec)
        pkv = V.mul(gen, sk)
        assert lib.h_verify_vargen(H.ptr(H.pt_mont(pkv)), H.ptr(H.pt_mont(gen)), H.ptr(H.limbs(eu)), H.ptr(H.pt_mont(eR)), H.ptr(H.mont(m)), 1, H.ptr(c)) == 1
        assert lib.h_verify_vargen(H.ptr(H.pt_mont(pkv, z1)), H.ptr(H.pt_mont(gen, z2)), H.ptr(H.limbs(eu)), H.ptr(H.pt_mont(eR, z1)), H.ptr(H.mont(m)), 0, H.ptr(c)) == 1
        assert lib.h_verify_vargen(H.ptr(H.pt_mont(pkv)), H.ptr(H.pt_mont(o.G)), H.ptr(H.limbs(eu)), H.ptr(H.pt_mont(eR)), H.ptr(H.mont(m)), 1, H.ptr(c)) == 0


def test_wire_formats(lib, tabs):
    """device-side JubJubAffine::{to,from}_bytes, from_bytes_wide and the byte-level verify, on the CPU build"""
    rnd = random.Random(8)
    out, uv, b32 = np.zeros(8, np.uint32), np.zeros(16, np.uint32), np.zeros(8, np.uint32)
    # sqrt: residues and non-residues, 0, 1, values whose x^t has maximal 2-power order
    for x in [0, 1, 4, Q - 1, 5, 7] + [rnd.randrange(Q) for _ in range(30)]:
        ok = lib.h_fq_sqrt(H.ptr(H.mont(x)), H.ptr(out))
        is_sq = x == 0 or pow(x, (Q - 1) // 2, Q) == 1
        assert bool(ok) == is_sq
        if is_sq:
            assert pow(H.unmont(out), 2, Q) == x
    # sqrt(num / den) by one exponentiation + table-driven discrete log: residues of every 2-power order class
    # (x = g^j z^2 ... covered by random x and by x with x^t of maximal order), non-residues, num = 0
    g = pow(7, (Q - 1) >> 32, Q)
    specials = [(0, 5), (1, 1), (4, 1), (1, 4), (Q - 1, 1), (g, 1), (pow(g, 2, Q), 1), (pow(g, 1 << 31, Q), 1), (pow(g, 6, Q), 9)]
    for n, d in specials + [(rnd.randrange(Q), rnd.randrange(1, Q)) for _ in range(60)]:
        ok = lib.h_fq_sqrt_ratio(H.ptr(H.mont(n)), H.ptr(H.mont(d)), H.ptr(out))
        x = n * pow(d, -1, Q) % Q
        is_sq = x == 0 or pow(x, (Q - 1) // 2, Q) == 1
        assert bool(ok) == is_sq, (n, d)
        if is_sq:
            assert pow(H.unmont(out), 2, Q) == x
    # decompress / compress
    pts = V.torsion_points() + [o.G, o.G_NUMS] + [V.rand_curve_point(rnd) for _ in range(10)]
    for P in pts:
        enc = o.affine_to_bytes(P)
        assert lib.h_decompress(H.ptr(np.frombuffer(enc, np.uint32).copy()), H.ptr(uv)) == 1
        assert (H.unmont(uv[:8]), H.unmont(uv[8:])) == P
        lib.h_compress(H.ptr(H.pt_mont(P)), H.ptr(b32))
        assert b32.tobytes() == enc
    for v in list(range(2, 40)) + [Q, Q + 1, (1 << 255) - 1]:
        enc = v.to_bytes(32, "little")
        exp = o.affine_from_bytes(enc)
        got = lib.h_decompress(H.ptr(np.frombuffer(enc, np.uint32).copy()), H.ptr(uv))
        assert bool(got) == (exp is not None)
        if exp is not None:
            assert (H.unmont(uv[:8]), H.unmont(uv[8:])) == exp
    # from_bytes_wide
    for w in [0, 1, R, Q, (1 << 512) - 1, (1 << 256) - 1, 1 << 256, R << 256] + [rnd.randrange(1 << 512) for _ in range(40)]:
        wb = np.frombuffer(w.to_bytes(64, "little"), np.uint32).copy()
        lib.h_fr_from_wide(H.ptr(wb), H.ptr(out)); assert H.to_int(out) == w % R
        lib.h_fq_from_wide(H.ptr(wb), H.ptr(out)); assert H.unmont(out) == w % Q
    # byte-level verify
    sk, nonce, m = rnd.randrange(R), rnd.randrange(R), rnd.randrange(Q)
    u, Rp, _ = o.sign(sk, nonce, m, mul=V.mul)
    pkb = o.affine_to_bytes(V.mul(o.G, sk))
    sig = u.to_bytes(32, "little") + o.affine_to_bytes(Rp)
    inv = ctypes.c_int(0)
    B = lambda b: H.ptr(np.frombuffer(b, np.uint32).copy())
    assert lib.h_verify_bytes(B(pkb), B(sig), B(m.to_bytes(32, "little")), H.ptr(tabs[0]), ctypes.byref(inv)) == 1 and inv.value == 0
    assert lib.h_verify_bytes(B(pkb), B(sig), B(((m + 1) % Q).to_bytes(32, "little")), H.ptr(tabs[0]), ctypes.byref(inv)) == 0 and inv.value == 0
    assert lib.h_verify_bytes(B(pkb), B(sig), B(Q.to_bytes(32, "little")), H.ptr(tabs[0]), ctypes.byref(inv)) == 0 and inv.value == 1
    bad_sig = (u + R).to_bytes(32, "little") + o.affine_to_bytes(Rp)
    assert lib.h_verify_bytes(B(pkb), B(bad_sig), B(m.to_bytes(32, "little")), H.ptr(tabs[0]), ctypes.byref(inv)) == 0 and inv.value == 1


# ---- FP64-pipe arithmetic (csrc/fd.cuh, hades_fd.cuh): 5 x 52-bit limbs, Montgomery radix 2^260 ----------------
R260INV = pow(1 << 260, -1, Q)


def test_fd_mul_sqr_mulc_add(lib):
    rnd = random.Random(11)
    out = np.zeros(8, np.uint32)
    top = (1 << 256) - 1  # lazily reduced operands (the test hooks carry 256-bit integers)
    vals = [0, 1, 2, Q - 1, Q, Q + 1, 2 * Q - 1, top, (1 << 255), (1 << 52) - 1, 1 << 52, ((1 << 52) - 1) << 52,
            sum(((1 << 52) - 1) << (52 * i) for i in range(4)) | (((1 << 48) - 1) << 208)] + [rnd.randrange(top) for _ in range(50)]
    for a in vals:
        lib.h_fd_sqr(H.ptr(H.limbs(a)), H.ptr(out))
        r = H.to_int(out)
        assert r % Q == a * a * R260INV % Q and r < a * a // (1 << 260) + Q + 1
        for b in vals[:20]:
            lib.h_fd_mul(H.ptr(H.limbs(a)), H.ptr(H.limbs(b)), H.ptr(out))
            r = H.to_int(out)
            assert r % Q == a * b * R260INV % Q and r < a * b // (1 << 260) + Q + 1
    # constant * x + running word, with the one conditional subtraction of q (word >= 2^255)
    topc = (1 << 256) - (1 << 251)
    for _ in range(300):
        cst, x = rnd.randrange(Q), rnd.randrange(2 * Q)
        c = rnd.choice([0, (1 << 255) - 1, 1 << 255, Q, topc, rnd.randrange(topc), rnd.randrange(Q)])  # result must fit the 256-bit hook
        lib.h_fd_mulc_add(H.ptr(H.limbs(cst)), H.ptr(H.limbs(x)), H.ptr(H.limbs(c)), H.ptr(out))
        r = H.to_int(out)
        cp = c - Q if c >= (1 << 255) else c
        assert r % Q == (cst * x * R260INV + c) % Q and cp <= r < cst * x // (1 << 260) + Q + 1 + cp


def test_fd_dot5_and_roundtrip(lib):
    rnd = random.Random(12)
    out = np.zeros(8, np.uint32)
    top = (1 << 256) - 1
    cases = [([Q - 1] * 5, [top] * 5, Q - 1), ([0] * 5, [top] * 5, 0), ([1, 0, 0, 0, 0], [5, 0, 0, 0, 0], None)]
    cases += [([rnd.randrange(Q) for _ in range(5)], [rnd.randrange(top) for _ in range(5)], rnd.choice([None, rnd.randrange(Q)])) for _ in range(200)]
    for c, s, add in cases:
        cb = np.concatenate([H.limbs(x) for x in c]); sb = np.concatenate([H.limbs(x) for x in s])
        lib.h_fd_dot5(H.ptr(cb), H.ptr(sb), H.ptr(H.limbs(add)) if add is not None else None, H.ptr(out))
        r = H.to_int(out)
        assert r % Q == (sum(a * b for a, b in zip(c, s)) * R260INV + (add or 0)) % Q
        assert r < sum(a * b for a, b in zip(c, s)) // (1 << 260) + Q + 1 + (add or 0)
    for a in [0, 1, Q - 1, Q - 2, 1 << 254] + [rnd.randrange(Q) for _ in range(100)]:
        lib.h_fd_roundtrip(H.ptr(H.mont(a)), H.ptr(out))
        assert H.to_int(out) == a


def test_hades_fd_matches_oracle(lib):
    rnd = random.Random(13)
    for st in [[0] * 5, [Q - 1] * 5, [1, 2, 3, 4, 5]] + [[rnd.randrange(Q) for _ in range(5)] for _ in range(6)]:
        buf = np.concatenate([H.mont(x) for x in st])
        lib.h_hades_fd(H.ptr(buf))
        assert [H.to_int(buf[8 * i:8 * i + 8]) for i in range(5)] == o.hades_perm(st)


def test_challenge_fd_matches_imad_path_and_oracle(lib):
    rnd = random.Random(14)
    c0, c1 = np.zeros(8, np.uint32), np.zeros(8, np.uint32)
    for _ in range(6):
        P1, P2 = V.mul(o.G, rnd.randrange(R)), V.mul(o.G_NUMS, rnd.randrange(R))
        m = rnd.choice([0, Q - 1, rnd.randrange(Q)])
        args = [H.ptr(H.mont(P1[0])), H.ptr(H.mont(P1[1])), H.ptr(H.mont(m))]
        lib.h_challenge3(*args, 0, H.ptr(c0)); lib.h_challenge3(*args, 1, H.ptr(c1))
        assert H.to_int(c0) == H.to_int(c1) == o.challenge_hash(P1, m)
        for mode in (2, 3):  # memory-operand permutation, dense and strided slot layouts
            lib.h_challenge3(*args, mode, H.ptr(c1))
            assert H.to_int(c1) == H.to_int(c0)
        args = [H.ptr(H.mont(P1[0])), H.ptr(H.mont(P1[1])), H.ptr(H.mont(P2[0])), H.ptr(H.mont(P2[1])), H.ptr(H.mont(m))]
        lib.h_challenge5(*args, 0, H.ptr(c0)); lib.h_challenge5(*args, 1, H.ptr(c1))
        assert H.to_int(c0) == H.to_int(c1) == o.challenge_hash_double(P1, P2, m)


def test_challenge_fd_pair(lib):
    """two interleaved permutations per call (the hash warps of the warp-specialised verify kernel)"""
    rnd = random.Random(15)
    out = np.zeros(16, np.uint32)
    for _ in range(4):
        P = [V.mul(o.G, rnd.randrange(R)) for _ in range(2)]
        m = [rnd.randrange(Q), rnd.choice([0, Q - 1])]
        ru = np.concatenate([H.mont(p[0]) for p in P]); rv = np.concatenate([H.mont(p[1]) for p in P])
        mm = np.concatenate([H.mont(x) for x in m])
        lib.h_challenge3_pair(H.ptr(ru), H.ptr(rv), H.ptr(mm), H.ptr(out))
        assert [H.to_int(out[:8]), H.to_int(out[8:])] == [o.challenge_hash(P[0], m[0]), o.challenge_hash(P[1], m[1])]


# ---- half-size scalars for verification (csrc/hgcd.cuh) --------------------------------------------------------
def test_half_gcd_invariants(lib):
    """a = b c (mod 8r), 0 <= a < 2^134, |b| < 2^134, b odd -- or ok = 0 (the caller then takes the full-size path)"""
    rnd = random.Random(21)
    N = 8 * R
    out = np.zeros(18, np.uint32)
    cases = [0, 1, 2, 3, (1 << 128) - 1, 1 << 128, (1 << 128) + 1, (1 << 250) - 1, N // 2 % (1 << 250), N // 3, (N // 3) * 2 % (1 << 250),
             (1 << 249), 8, 16, R % (1 << 250), (N - 1) % (1 << 250)]
    cases += [rnd.randrange(1 << 250) for _ in range(400)]
    cases += [rnd.randrange(1 << 130) for _ in range(50)]               # barely above the bound: few steps
    cases += [pow(rnd.randrange(1, 1 << 20), -1, N) % (1 << 250) if False else (N // rnd.randrange(2, 1 << 40)) % (1 << 250) for _ in range(100)]  # N / small: huge first quotients
    n_ok = 0
    for c in cases:
        lib.h_half_gcd(H.ptr(H.limbs(c)), H.ptr(out))
        a, b = H.to_int(out[:8]), H.to_int(out[8:16])
        bneg, ok = int(out[16]), int(out[17])
        if not ok:
            continue
        n_ok += 1
        bs = -b if bneg else b
        assert a < (1 << 134) and b < (1 << 134) and b % 2 == 1, (c, a, b)
        assert (a - bs * c) % N == 0, (c, a, bs)
    assert n_ok >= len(cases) - 110  # the window budget is missed only by pathological c (the 100 N / small cases)


def test_verify_half_size_equals_full_size(lib, tabs):
    """the half-size predicate gives the reference's verdict on the whole curve: valid and corrupted signatures,
    keys / nonce points with an 8-torsion component, identity and small-order keys"""
    import ctypes
    rnd = random.Random(22)
    fo = ctypes.c_int(0)
    tors = V.torsion_points()
    def both(pk, u, Rp, c):
        args = [H.ptr(H.pt_mont(pk)), H.ptr(H.limbs(u)), H.ptr(H.pt_mont(Rp)), H.ptr(H.limbs(c)), 1, H.ptr(tabs[0])]
        slow = lib.h_verify_ec(*args, 0, ctypes.byref(fo))
        fast = lib.h_verify_ec(*args, 1, ctypes.byref(fo))
        assert fo.value == 1
        return slow, fast
    for trial in range(10):
        sk, nonce, m = rnd.randrange(R), rnd.randrange(R), rnd.randrange(Q)
        u, Rp, c = o.sign(sk, nonce, m, mul=V.mul)
        pk = V.mul(o.G, sk)
        assert both(pk, u, Rp, c) == (1, 1)
        assert both(pk, (u + 1) % R, Rp, c) == (0, 0)
        assert both(pk, u, o.pt_add(Rp, o.G), c) == (0, 0)
        # torsion: PK' = PK + T, R' = R + c T verifies under the reference's equation (exact integer multiples)
        T = tors[trial % len(tors)]
        pkt = o.pt_add(pk, T)
        Rt = o.pt_add(Rp, V.mul(T, c))
        exp = int(o.pt_add(V.mul(o.G, u), V.mul(pkt, c)) == Rt)
        assert exp == 1 and both(pkt, u, Rt, c) == (1, 1)
        # ... and R' = R + c T + (another torsion point) does not
        T2 = tors[(trial + 3) % len(tors)]
        if T2 != o.IDENTITY:
            assert both(pkt, u, o.pt_add(Rt, T2), c) == (0, 0)
        # small-order key alone: accepts iff u G + c T == R
        Rs = o.pt_add(V.mul(o.G, u), V.mul(T, c))
        assert both(T, u, Rs, c) == (1, 1)


def test_hostile_challenges_defeat_half_gcd_and_fallback_is_exact(lib):
    """tests/vectors.py hgcd_hostile_challenges: the half-size path reports ok = 0 for them, and the full-size path the
    kernel falls back to gives the oracle's verdict (the GPU twin: tests/test_gpu_robustness.py)"""
    import ctypes
    rnd = random.Random(77)
    tabs = H.comb_tables(lib)
    for c in V.hgcd_hostile_challenges(4):
        out = np.zeros(18, np.uint32)
        lib.h_half_gcd(H.ptr(H.limbs(c)), H.ptr(out))
        assert out[17] == 0 and c < (1 << 250)
        P, u = V.mul(o.G, rnd.randrange(R)), rnd.randrange(R)
        good = o.pt_add(V.mul(o.G, u), V.mul(P, c))
        fo = ctypes.c_int(1)
        for Rp, want in ((good, 1), (o.pt_add(good, o.G), 0)):
            got = lib.h_verify_ec(H.ptr(H.pt_mont(P)), H.ptr(H.limbs(u)), H.ptr(H.pt_mont(Rp)), H.ptr(H.limbs(c)), 1, H.ptr(tabs[0]), 0,
                                  ctypes.byref(fo))
            assert got == want
        lib.h_verify_ec(H.ptr(H.pt_mont(P)), H.ptr(H.limbs(u)), H.ptr(H.pt_mont(good)), H.ptr(H.limbs(c)), 1, H.ptr(tabs[0]), 1, ctypes.byref(fo))
        assert fo.value == 0


def test_point_well_formed(lib):
    rnd = random.Random(78)
    for _ in range(6):
        P = V.rand_curve_point(rnd)
        z = rnd.randrange(1, Q)
        assert lib.h_point_well_formed(H.ptr(H.pt_mont(P)), 1) == 1
        assert lib.h_point_well_formed(H.ptr(H.pt_mont(P, z)), 0) == 1
        assert lib.h_point_well_formed(H.ptr(H.pt_mont((P[0], (P[1] + 1) % Q))), 1) == 0
        bad = H.pt_mont(P, z)
        bad[16:24] = 0
        assert lib.h_point_well_formed(H.ptr(bad), 0) == 0


@pytest.mark.parametrize("scheme", [0, 1, 2])
def test_witness_rows_host_build(lib, scheme):
    """core.cuh witness_core on the CPU build: u, R, PK, m, c, SA = u G, SB = c PK as the gadgets allocate them"""
    rnd = random.Random(90 + scheme)
    tabs = H.comb_tables(lib)
    sk, nonce, m = rnd.randrange(1, R), rnd.randrange(R), rnd.randrange(Q)
    gen, z = V.mul(o.G, rnd.randrange(1, R)), rnd.randrange(1, Q)
    w = (11, 19, 13)[scheme]
    rows = np.zeros(8 * w, np.uint32)
    lib.h_witness(scheme, H.ptr(H.limbs(sk)), H.ptr(H.limbs(nonce)), H.ptr(H.mont(m)), H.ptr(H.pt_mont(gen, z)), 0,
                  H.ptr(tabs[0]), H.ptr(tabs[1]), H.ptr(rows))
    got = [H.unmont(rows[8 * k:8 * k + 8]) for k in range(w)]
    if scheme == 0:
        u, Rp, c = o.sign(sk, nonce, m, mul=V.mul)
        pk = V.mul(o.G, sk)
        want = [u, *Rp, *pk, m, c, *V.mul(o.G, u), *V.mul(pk, c)]
    elif scheme == 1:
        u, Rp, Rpp, c = o.sign_double(sk, nonce, m, mul=V.mul)
        pk, pkp = V.mul(o.G, sk), V.mul(o.G_NUMS, sk)
        want = [u, *Rp, *Rpp, *pk, *pkp, m, c, *V.mul(o.G, u), *V.mul(pk, c), *V.mul(o.G_NUMS, u), *V.mul(pkp, c)]
    else:
        u, Rp, c = o.sign_vargen(sk, gen, nonce, m, mul=V.mul)
        pk = V.mul(gen, sk)
        want = [u, *Rp, *pk, *gen, m, c, *V.mul(gen, u), *V.mul(pk, c)]
    assert got == want


def test_oblivious_scalar_multiplication(lib):
    """SB200_SIGN_OBLIVIOUS paths (ed.cuh): 4-bit comb read by masked scan / scanned window table give the oracle's
    multiples for edge and random scalars"""
    rnd = random.Random(91)
    tab = np.zeros(64 * 8 * 24, np.uint32)
    lib.h_comb4_build(H.ptr(H.mont(o.G[0])), H.ptr(H.mont(o.G[1])), H.ptr(tab))
    out = np.zeros(16, np.uint32)
    for k in [0, 1, 7, 8, 9, 15, 16, 0x88, R - 1, R, (1 << 252) - 1] + [rnd.randrange(R) for _ in range(6)]:
        lib.h_fixed_mul_oblivious(H.ptr(tab), H.ptr(H.limbs(k)), H.ptr(out))
        assert (H.unmont(out[:8]), H.unmont(out[8:])) == V.mul(o.G, k), hex(k)
    P = V.rand_curve_point(rnd)  # whole curve: may carry a torsion component
    for k in [0, 1, 8, R - 1, (1 << 252) - 1, rnd.randrange(R)]:
        for z in (None, rnd.randrange(1, Q)):
            lib.h_var_mul_oblivious(H.ptr(H.pt_mont(P, z)), int(z is None), H.ptr(H.limbs(k)), H.ptr(out))
            assert (H.unmont(out[:8]), H.unmont(out[8:])) == V.mul(P, k), hex(k)


def _lat3(lib, c, u):
    out = np.zeros(28, np.uint32)
    lib.h_lattice3(H.ptr(H.limbs(c)), H.ptr(H.limbs(u)), H.ptr(out))
    a, b, d = H.to_int(out[:8]), H.to_int(out[8:16]), H.to_int(out[16:24])
    return (-a if out[24] else a), (-b if out[25] else b), (-d if out[26] else d), bool(out[27])


def test_lattice3_invariants(lib):
    """csrc/lat3.cuh: (b, a, d) with a = b c, d = b u (mod 8r), b odd, all below the 174-bit window budget --
    for edge and random (c, u); `ok` may only be false if the budget is missed, never true on a wrong vector"""
    rnd = random.Random(95)
    N = 8 * R
    edge = [0, 1, 2, R - 1, (1 << 250) - 1, N // 3 % (1 << 250), (1 << 128), (1 << 170) + 1]
    cases = [(c, u) for c in edge for u in (0, 1, R - 1, rnd.randrange(R))] + [(rnd.randrange(1 << 250), rnd.randrange(R)) for _ in range(300)]
    cases += [(c, rnd.randrange(R)) for c in V.hgcd_hostile_challenges(4)]
    n_ok_random = 0
    for k, (c, u) in enumerate(cases):
        a, b, d, ok = _lat3(lib, c, u)
        if not ok:  # e.g. u = r - 1: every short vector has an even b ((8, ., -8) is in the lattice); the kernel falls back
            continue
        n_ok_random += 4 * len(edge) <= k < 4 * len(edge) + 300
        assert b % 2 == 1 and b != 0
        assert (a - b * c) % N == 0 and (d - b * u) % N == 0, (hex(c), hex(u))
        assert max(abs(a), abs(b), abs(d)) < (1 << 174)
    assert n_ok_random >= 299  # random inputs essentially always meet the budget


def test_verify_vargen_short_scalars_equal_full_size(lib):
    """the 3-table short-scalar form gives the verdict of the reference-shaped full-size form: valid, corrupted,
    keys / generators / nonce points with torsion components, projective inputs"""
    import ctypes
    rnd = random.Random(96)
    tors = V.torsion_points()
    fo = ctypes.c_int(1)
    for trial in range(10):
        gen = V.mul(o.G, rnd.randrange(1, R))
        sk, nonce, m = rnd.randrange(R), rnd.randrange(R), rnd.randrange(Q)
        if trial % 3 == 1:
            gen = o.pt_add(gen, tors[trial % len(tors)])   # generator outside the prime-order subgroup
        pk = V.mul(gen, sk)
        if trial % 3 == 2:
            pk = o.pt_add(pk, tors[(trial + 3) % len(tors)])
        c = rnd.randrange(1 << 250)
        u = rnd.randrange(R)
        good = o.pt_add(V.mul(gen, u), V.mul(pk, c))      # R that makes u Gen + c PK == R hold
        for Rp, want in ((good, 1), (o.pt_add(good, o.G), 0), (o.pt_add(good, tors[1]), 0)):
            z = rnd.randrange(1, Q)
            for aff, zz in ((1, None), (0, z)):
                args = (H.ptr(H.pt_mont(pk, zz)), H.ptr(H.pt_mont(gen, zz)), H.ptr(H.limbs(u)), H.ptr(H.pt_mont(Rp, zz)), H.ptr(H.limbs(c)), aff)
                assert lib.h_verify_vargen_ec(*args, 1, ctypes.byref(fo)) == want and fo.value == 1
                assert lib.h_verify_vargen_ec(*args, 0, ctypes.byref(fo)) == want
    # u >= r is rejected by both
    args = (H.ptr(H.pt_mont(pk)), H.ptr(H.pt_mont(gen)), H.ptr(H.limbs(R)), H.ptr(H.pt_mont(good)), H.ptr(H.limbs(c)), 1)
    assert lib.h_verify_vargen_ec(*args, 1, ctypes.byref(fo)) == 0 and lib.h_verify_vargen_ec(*args, 0, ctypes.byref(fo)) == 0


def test_euclidean_inversion_equals_fermat(lib):
    """csrc/inv.cuh: the two-level Euclidean inversion returns exactly a^-1 (Montgomery in, Montgomery out) -- against the big-int
    inverse and against the Fermat form it replaces -- for edge values (0 -> 0, 1, q - 1, 2, 2^255 mod q, values whose Euclid has
    huge / tiny quotients) and random ones"""
    rnd = random.Random(98)
    vals = [0, 1, 2, 3, Q - 1, Q - 2, (Q - 1) // 2, (Q + 1) // 2, 1 << 128, (1 << 128) - 1, (1 << 254) % Q, pow(2, 256, Q), pow(3, Q - 2, Q),
            Q // 3, Q // 65537, (Q // 3) * 2 + 1, 0xFFFFFFFF, 1 << 32, (1 << 200) + 1]
    vals += [rnd.randrange(Q) for _ in range(400)]
    vals += [rnd.randrange(1 << rnd.randrange(1, 255)) for _ in range(100)]
    for a in vals:
        out, ref = np.zeros(8, np.uint32), np.zeros(8, np.uint32)
        lib.h_fq_inv_fast(H.ptr(H.mont(a)), H.ptr(out))
        lib.h_fq_inv_fermat(H.ptr(H.mont(a)), H.ptr(ref))
        want = pow(a, Q - 2, Q)
        assert H.unmont(out) == want, hex(a)
        assert (out == ref).all()


def test_hgcd_double_precision_quotient_never_overshoots():
    """csrc/hgcd.cuh estimates the Euclid's partial quotient as floor(double(R0) / (double(R1) + 1) * (1 - 2^-50)) from the top 64
    bits of the two remainders.  The step is only valid if that never exceeds floor(R0 / (R1 + 1)) (<= floor(r0 / r1)): checked
    here in IEEE double arithmetic (Python floats) on random and adversarial 64-bit pairs -- exact multiples, neighbours of
    powers of two, values whose conversion to double rounds up."""
    rnd = random.Random(99)
    k = 0.99999999999999911182158029987
    assert k == 1.0 - 2.0 ** -50
    pairs = []
    for _ in range(20000):
        r1 = rnd.getrandbits(rnd.randrange(1, 65))
        q = rnd.getrandbits(rnd.randrange(1, 33))
        pairs.append((min(q * (r1 + 1) + rnd.choice([0, 0, 1, r1]), (1 << 64) - 1), r1))   # at or just above an exact multiple
        pairs.append((rnd.getrandbits(64), rnd.getrandbits(rnd.randrange(1, 65))))
    for e in range(1, 64):
        for d in (-1, 0, 1):
            pairs += [((1 << 64) - 1, (1 << e) + d), ((1 << 63) + d + 1, (1 << e) + d), ((1 << 53) + (1 << e) + d, (1 << (e % 53)) + 1)]
    pairs += [((1 << 64) - 1, 0), ((1 << 64) - 1, (1 << 64) - 1), ((1 << 64) - 1, (1 << 64) - 2), (1, 0), (0, 5)]
    for R0, R1 in pairs:
        if R1 < 0:
            continue
        est = int(float(R0) / (float(R1) + 1.0) * k)
        true = R0 // (R1 + 1)
        assert est <= true, (R0, R1, est, true)
        # an underestimate only costs an extra step; below the clamp (2^31 - 1) it is at most one
        assert est >= min(true, (1 << 31) - 1) - 1, (R0, R1, est, true)
