"""The N>1 path on the CPU: two gloo ranks shard a batch by tuple index exactly like bench.py / the C library
do (contiguous blocks, multiples of 32, no data-path collective), each produces its shard's verdict words
(with the CPU checker standing in for the kernel), and the merged bitmap equals the single-process result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sharding import bitmap_words, shard_range


def test_shard_ranges_cover_and_align():
    for n in (0, 1, 31, 32, 33, 1000, 1 << 16, (1 << 16) + 5):
        for w in (1, 2, 4, 8):
            rs = [shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert all(lo % 32 == 0 or lo == hi == n for lo, hi in rs)  # only empty tail shards start unaligned


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "oracle"), os.path.join(root, "tests")]
    import ref_cpu
    rs = np.random.RandomState(7)
    sk = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32); sk[:, 7] &= (1 << 27) - 1
    nonce = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32); nonce[:, 7] &= (1 << 27) - 1
    msg = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32); msg[:, 7] &= (1 << 30) - 1
    lo, hi = shard_range(n, rank, world)
    pk = ref_cpu.keygen(sk[lo:hi], threads=1)
    u, R, _ = ref_cpu.sign(sk[lo:hi], msg[lo:hi], nonce[lo:hi], threads=1)
    u[(np.arange(lo, hi) % 5) == 0, 0] ^= 1
    ok, _ = ref_cpu.verify(pk, u, R, msg[lo:hi], threads=1)
    words = np.packbits(np.concatenate([ok, np.zeros((-len(ok)) % 32, bool)]), bitorder="little").view(np.uint32)
    a, b = bitmap_words(lo, hi)
    assert len(words) == b - a
    # timing contract of bench.py: every rank contributes its time, the MAX is reported
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    # verdict words are only gathered here to CHECK the shards; the product path has no collective
    full = torch.zeros((n + 31) // 32, dtype=torch.int64)
    full[a:b] = torch.from_numpy(words.astype(np.int64))
    dist.all_reduce(full)
    if rank == 0:
        q.put(full.numpy().astype(np.uint32))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_verify_matches_single_process():
    n, world = 150, 2  # ragged: 150 = 96 + 54
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    merged = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ok = np.unpackbits(merged.view(np.uint8), bitorder="little")[:n].astype(bool)
    assert ok.tolist() == [(i % 5) != 0 for i in range(n)]
