#!/usr/bin/env python3
"""Strong scaling of ONE batch through ONE context: `Engine(list(range(N)))` splits a single batch of
2^22 single-key verifications held in host buffers over N = 1, 2, 4, 8 devices (one host thread per device, no
collective, no torchrun).  Prints one JSON line per N; the limiter is named from the measured components.
usage: tools/strong_scaling.py [log2n] [--pageable]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from schnorr_b200 import POINTS_AFFINE, Engine, PinnedBuffer  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 22
pageable = "--pageable" in sys.argv
n = 1 << lg
ndev = torch.cuda.device_count()
rs = np.random.RandomState(0xC1)


def sc(bits):
    a = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    a[:, 7] &= (1 << bits) - 1
    return a


sk, nonce, msg = sc(27), sc(27), sc(30)
e1 = Engine([0])
pk = e1.keygen(sk)
u, R, _ = e1.sign(sk, msg, nonce)
bad = (np.arange(n, dtype=np.int64) * 2654435761 % 10) == 0
u[bad, 0] ^= 1
e1.close()
if pageable:
    bufs = [np.ascontiguousarray(a) for a in (pk, u, R, msg)] + [np.zeros((n + 31) // 32 + 16, np.uint32)]
    off = (-bufs[4].ctypes.data // 4) % 4
    bufs[4] = bufs[4][off:off + (n + 31) // 32]
    keep = None
else:
    keep = [PinnedBuffer(s) for s in ((n, 16), (n, 8), (n, 16), (n, 8), ((n + 31) // 32,))]
    bufs = [b.array for b in keep]
    for dst, src in zip(bufs[:4], (pk, u, R, msg)):
        dst[...] = src
P = lambda a: a.ctypes.data
h2d = n * 192
base = None
for N in (1, 2, 4, 8):
    if N > ndev:
        break
    eng = Engine(list(range(N)))
    call = lambda: eng.call("verify", n, POINTS_AFFINE, P(bufs[0]), P(bufs[1]), P(bufs[2]), P(bufs[3]), P(bufs[4]), None)
    call(); call()
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        call()
    dt = (time.perf_counter() - t0) / reps
    ok = np.unpackbits(bufs[4].view(np.uint8), bitorder="little")[:n].astype(bool)
    assert (ok == ~bad).all(), "verdicts wrong"
    base = base or dt
    rec = {"n_gpus": N, "n_total": n, "verifies_per_s": n / dt, "ms_per_batch": dt * 1e3, "speedup_vs_1": base / dt,
           "efficiency": base / dt / N, "h2d_gbs_total": h2d / dt / 1e9, "h2d_gbs_per_gpu": h2d / dt / 1e9 / N,
           "host_memory": "pageable (staged by the library)" if pageable else "pinned", "verdicts": "all correct",
           "chunks_per_device": max(4, (n // N + (1 << 18) - 1) >> 18)}
    print(json.dumps(rec), flush=True)
    eng.close()
