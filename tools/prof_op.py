"""Tiny device-resident run of one op for ncu (development / profiling aid).  usage: prof_op.py <op> <log2n> [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from schnorr_b200 import Engine, POINTS_AFFINE, DEVICE_PTRS
op = sys.argv[1] if len(sys.argv) > 1 else "verify"
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 17)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
e = Engine([0])
rs = np.random.RandomState(1)
def sc(bits):
    a = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32); a[:, 7] &= (1 << bits) - 1; return a
sk, nonce, msg = sc(27), sc(27), sc(30)
pk = e.keygen(sk); u, R, _ = e.sign(sk, msg, nonce)
dev = torch.device("cuda:0")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(dev)
d_pk, d_u, d_R, d_msg, d_sk, d_nonce = t(pk), t(u), t(R), t(msg), t(sk), t(nonce)
bm = torch.zeros((n + 31) // 32, dtype=torch.int32, device=dev)
uo, Ro, Ro2, co = (torch.empty((n, k), dtype=torch.int32, device=dev) for k in (8, 16, 16, 8))
e.set_stream(torch.cuda.current_stream().cuda_stream)
fl = POINTS_AFFINE | DEVICE_PTRS
vfl = fl | (4 if os.environ.get("SB_DUAL") else 0)  # SB200_VERIFY_DUAL_PIPE
P = lambda x: x.data_ptr()
calls = {
  "verify": lambda: e.call("verify", n, vfl, P(d_pk), P(d_u), P(d_R), P(d_msg), P(bm), P(co) if os.environ.get("SB_COUT") else None),
  "verify_vargen": lambda: e.call("verify_vargen", n, fl, P(d_pk), P(d_pk), P(d_u), P(d_R), P(d_msg), P(bm), None),
  "verify_double": lambda: e.call("verify_double", n, fl, P(d_pk), P(d_pk), P(d_u), P(d_R), P(d_R), P(d_msg), P(bm), None),
  "sign": lambda: e.call("sign", n, fl, P(d_sk), P(d_msg), P(d_nonce), P(uo), P(Ro), P(co)),
  "sign_double": lambda: e.call("sign_double", n, fl, P(d_sk), P(d_msg), P(d_nonce), P(uo), P(Ro), P(Ro2), P(co)),
}
s, f = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
calls[op](); torch.cuda.synchronize()
s.record()
for _ in range(reps): calls[op]()
f.record(); torch.cuda.synchronize()
ms = s.elapsed_time(f) / reps
if op == "verify":
    ok = np.unpackbits(bm.cpu().numpy().view(np.uint8), bitorder="little")[:n]
    assert ok.all(), "valid signatures must verify"
print(f"{op} n={n} {ms:.3f} ms/launch {n/ms/1e3:.3f} M/s")
