"""Scratch timing of the kernels with device-resident inputs (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from schnorr_b200 import Engine, POINTS_AFFINE, DEVICE_PTRS

e = Engine([0])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
rs = np.random.RandomState(1)
# synthetic well-formed-looking limbs (< q, < r): timing does not depend on validity
def rnd_fq(n, words=8):
    a = rs.randint(0, 1 << 32, size=(n, words), dtype=np.uint64).astype(np.uint32)
    a[:, 7::8] &= 0x0FFFFFFF
    return a
dev = torch.device("cuda:0")
t = lambda a: torch.from_numpy(a.view(np.int32)).to(dev)
pk, u, R, m = t(rnd_fq(n, 16)), t(rnd_fq(n, 8)), t(rnd_fq(n, 16)), t(rnd_fq(n, 8))
pk2, R2 = t(rnd_fq(n, 16)), t(rnd_fq(n, 16))
bm = torch.zeros((n + 31) // 32, dtype=torch.int32, device=dev)
uo, Ro, Ro2, co = t(np.zeros((n, 8), np.uint32)), t(np.zeros((n, 16), np.uint32)), t(np.zeros((n, 16), np.uint32)), t(np.zeros((n, 8), np.uint32))
e.set_stream(torch.cuda.current_stream().cuda_stream)
fl = POINTS_AFFINE | DEVICE_PTRS
def timeit(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    s, f = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    f.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(f) / reps
    print(f"{name:16s} n={n} {ms:9.3f} ms  {n / ms / 1e3:8.3f} M/s", flush=True)
P = lambda x: x.data_ptr()
timeit("verify", lambda: e.call("verify", n, fl, P(pk), P(u), P(R), P(m), P(bm), None))
timeit("verify_vargen", lambda: e.call("verify_vargen", n, fl, P(pk), P(pk2), P(u), P(R), P(m), P(bm), None))
timeit("verify_double", lambda: e.call("verify_double", n, fl, P(pk), P(pk2), P(u), P(R), P(R2), P(m), P(bm), None))
timeit("sign", lambda: e.call("sign", n, fl, P(u), P(m), P(u), P(uo), P(Ro), P(co)))
timeit("sign_double", lambda: e.call("sign_double", n, fl, P(u), P(m), P(u), P(uo), P(Ro), P(Ro2), P(co)))
timeit("keygen", lambda: e.call("keygen", n, fl, P(u), P(Ro)))

# byte-level entry points (valid inputs needed: decompression must succeed)
nb = min(n, 1 << 18)
sk_np = rnd_fq(nb, 8); sk_np[:, 7] &= 0x07FFFFFF
msg_np = rnd_fq(nb, 8)
sigb = e.sign_bytes(sk_np.view(np.uint8), msg_np.view(np.uint8), sk_np.view(np.uint8))
pkb = e.points_compress(e.keygen(sk_np))
tb = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(dev)
d_pkb, d_sigb, d_msgb, d_skb = tb(pkb), tb(sigb), tb(msg_np), tb(sk_np)
d_sigo = torch.empty((nb, 16), dtype=torch.int32, device=dev)
d_pts = torch.empty((nb, 16), dtype=torch.int32, device=dev)
bm2 = torch.zeros((nb + 31) // 32, dtype=torch.int32, device=dev)
n_save, n = n, nb
timeit("verify_bytes", lambda: e.call("verify_bytes", nb, DEVICE_PTRS, P(d_pkb), P(d_sigb), P(d_msgb), P(bm2), None))
import numpy as _np
assert _np.unpackbits(bm2.cpu().numpy().view(_np.uint8), bitorder="little")[:nb].all(), "byte-level verify of valid signatures"
timeit("sign_bytes", lambda: e.call("sign_bytes", nb, DEVICE_PTRS, P(d_skb), P(d_msgb), P(d_skb), P(d_sigo), None))
timeit("decompress", lambda: e.call("points_decompress", nb, DEVICE_PTRS, P(d_pkb), P(d_pts), P(bm2)))
timeit("compress", lambda: e.call("points_compress", nb, fl, P(d_pts), P(d_pkb)))
