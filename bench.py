#!/usr/bin/env python3
"""bench.py -- the headline measurement of the sign/verify hot path.

    python bench.py --gpus N --steps K --warmup W            (N>1: under torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload): single-key `PublicKey::verify` (/root/reference/src/keys/public.rs:121-130)
over a batch of 2^22 (message, key, signature) tuples PER GPU (weak scaling; tuples are independent,
sharded by index, no collective), 10 % of the signatures corrupted.  A "step" = one pass of the
verify path over the whole batch.  Inputs are synthetic: seeded keys / messages / nonces, signed with
the engine's own batch signer (whose parity with the oracle is what tests/ establish).

  value  = verifies/s, inputs resident in HBM (DEVICE_PTRS call, CUDA events on the launching stream)
  e2e    = verifies/s through the host-buffer C-ABI call (pinned host inputs, H2D + kernel + D2H of
           the verdict bitmap inside the timed region)
  roofline = IMAD.WIDE issue roofline (the path is integer-multiply bound; HBM share reported too)
  cpu_baseline = oracle/ref_cpu.c (C restatement of the reference's algorithm, all host cores) on a
           bounded prefix of the same batch; its verdicts are also compared with the GPU's.

`--impl reference` times that CPU restatement alone (the Rust crate cannot be built in this image).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG2_BATCH = 22
METRIC = "verifies_per_s"
UNIT = "verifies/s"
# 32x32->64 limb products one single-key verification executes in k_run<OP_VERIFY> (affine inputs),
# counted by the instrumented host build of the same code (tests/test_op_counts.py keeps this honest).
OPCOUNT_FILE = os.path.join(ROOT, "profiles", "op_counts.json")
IMAD_PEAK_FILE = os.path.join(ROOT, "IMAD_PEAK.json")


def synth_scalars(rs, n, top_bits):
    a = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    a[:, 7] &= (1 << top_bits) - 1
    return a


def make_inputs(n, seed):
    """sk, nonce < 2^251 < r (canonical scalars); msg < 2^254 < q (Montgomery limbs of some field element)."""
    rs = np.random.RandomState(seed)
    return synth_scalars(rs, n, 27), synth_scalars(rs, n, 27), synth_scalars(rs, n, 30)


def corrupt_mask(n):
    return (np.arange(n, dtype=np.int64) * 2654435761 % 10) == 0  # deterministic ~10 %


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz, self.power = index, [], set(), False, None, []
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {"GpuIdle": "gpu_idle", "SwPowerCap": "sw_power_cap", "HwSlowdown": "hw_slowdown",
                 "SwThermalSlowdown": "sw_thermal_slowdown", "HwThermalSlowdown": "hw_thermal_slowdown",
                 "HwPowerBrakeSlowdown": "hw_power_brake", "ApplicationsClocksSetting": "applications_clocks_setting",
                 "SyncBoost": "sync_boost", "DisplayClockSetting": "display_clock_setting"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, v in names.items():
                    bit = getattr(nv, "nvmlClocksThrottleReason" + k, 0)
                    if bit and (r & bit):
                        self.reasons.add(v)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(x for x in self.reasons if x != "gpu_idle"), "samples": len(s),
                "power_w_max": max(self.power) if self.power else None}


def ncu_traffic(n):
    """dram__bytes_read.sum + dram__bytes_write.sum of the verify kernel from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by tools/ncu_traffic.py), scaled to this launch's tuple count; None if absent."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["verify_affine"]
        return (t["dram_bytes_read"] + t["dram_bytes_write"]) * n / t["tuples"]
    except Exception:
        return None


def cpu_reference_arm(args):
    """The reference's CPU implementation of the path = oracle/ref_cpu.c, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_cpu
    cores = os.cpu_count() or 1
    n = args.ref_sample or 2048 * cores
    sk, nonce, msg = make_inputs(n, 0xC1)
    pk = ref_cpu.keygen(sk)
    u, R, _ = ref_cpu.sign(sk, msg, nonce)
    bad = corrupt_mask(n)
    u[bad, 0] ^= 1
    for _ in range(args.warmup):
        ref_cpu.verify(pk[:256 * cores], u[:256 * cores], R[:256 * cores], msg[:256 * cores])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ok, _ = ref_cpu.verify(pk, u, R, msg)
    dt = (time.perf_counter() - t0) / args.steps
    assert (ok == ~bad).all(), "CPU restatement verdicts are wrong"
    val = n / dt
    sample = f"{n} tuples/step (prefix of the 2^{LOG2_BATCH} workload), {cores} pthreads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"single-key PublicKey::verify, 10% corrupted, CPU restatement of dusk-schnorr's algorithm "
                               f"(the Rust crate cannot be built in this image), bounded sample of {n} tuples per step"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-batch", type=int, default=LOG2_BATCH)
    ap.add_argument("--ref-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return cpu_reference_arm(args)

    import torch
    import torch.distributed as dist
    from schnorr_b200 import DEVICE_PTRS, POINTS_AFFINE, VERIFY_DUAL_PIPE, Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # NCCL_DEBUG=VERSION (some boxes export it) makes NCCL print its banner on stdout, next to the one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    eng = Engine([local])  # raises without a GPU: there is no CPU fallback
    n = 1 << args.log2_batch

    # ---- synthetic workload, built with the engine's own signer ---------------------------------
    sk, nonce, msg = make_inputs(n, 0xC1 + rank)
    pinned = lambda shape: torch.empty(shape, dtype=torch.int32, pin_memory=True)
    h_pk, h_u, h_R, h_msg = pinned((n, 16)), pinned((n, 8)), pinned((n, 16)), pinned((n, 8))
    h_bm = pinned(((n + 31) // 32,))
    npv = lambda t: t.numpy().view(np.uint32)
    pk = eng.keygen(sk)
    u, R, _ = eng.sign(sk, msg, nonce)
    bad = corrupt_mask(n)
    u[bad, 0] ^= 1  # still canonical (< r): flips the lowest bit
    npv(h_pk)[...] = pk; npv(h_u)[...] = u; npv(h_R)[...] = R; npv(h_msg)[...] = msg
    d_pk, d_u, d_R, d_msg = (t.to(dev) for t in (h_pk, h_u, h_R, h_msg))
    d_bm = torch.zeros((n + 31) // 32, dtype=torch.int32, device=dev)
    P = lambda t: t.data_ptr()
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    fl = POINTS_AFFINE | DEVICE_PTRS

    def step_dev():
        eng.call("verify", n, fl, P(d_pk), P(d_u), P(d_R), P(d_msg), P(d_bm), None)

    def step_e2e():
        eng.call("verify", n, POINTS_AFFINE, P(h_pk), P(h_u), P(h_R), P(h_msg), P(h_bm), None)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: device-resident, CUDA events on the launching stream ------------------------------
    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(local)
    barrier()
    launches0 = eng.launch_count
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
    e1.record(stream)
    barrier()
    sampler.stop_flag = True
    launches = eng.launch_count - launches0
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    verdict = np.unpackbits(d_bm.cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
    assert (verdict == ~bad).all(), "GPU verdicts wrong: valid signatures must verify, corrupted ones must not"
    value = world * n / (ms_step * 1e-3)

    # ---- e2e: host buffers through the C ABI ---------------------------------------------------------
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    assert (np.unpackbits(npv(h_bm).view(np.uint8), bitorder="little")[:n].astype(bool) == ~bad).all()
    e2e = world * n / e2e_s
    h2d = n * (16 + 8 + 16 + 8) * 4
    d2h = ((n + 31) // 32) * 4

    # ---- roofline of the dominant (only) kernel ---------------------------------------------------------
    roof = None
    try:
        peak = json.load(open(IMAD_PEAK_FILE))
        ops = json.load(open(OPCOUNT_FILE))["verify_affine"]
        wide = ops["imad_wide_per_tuple"]
        achieved = n * wide / (ms_step * 1e-3) / 1e12
        roof = {"bound": "imad", "achieved": achieved, "peak": peak["imad_wide_per_s"] / 1e12, "unit": "T IMAD.WIDE/s",
                "frac": achieved / (peak["imad_wide_per_s"] / 1e12), "traffic": ncu_traffic(n),
                "peak_source": "IMAD_PEAK.json (tools/imad_peak.cu measured on this pool's B200: IMAD.WIDE carry-chain issue rate)",
                "work": f"{wide} IMAD.WIDE.U32 per verification ({ops['fq_mul']} fq_mul, {ops['fq_sqr']} fq_sqr, {ops.get('fq_dot5', 0)} fq_dot5, "
                        f"{ops.get('fr_mont_mul', 0)} fr_mont_mul, the rest the half-size-scalar Euclid)",
                "hbm": {"achieved_gbs": (h2d + d2h) / (ms_step * 1e-3) / 1e9,
                        "peak_gbs": json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
                        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0}}
        roof["hbm"]["frac"] = roof["hbm"]["achieved_gbs"] / roof["hbm"]["peak_gbs"]
    except Exception as ex:  # pragma: no cover
        roof = {"bound": "imad", "error": repr(ex)}

    # ---- the other ops of the path (informational) ---------------------------------------------------------
    extras = {}
    if not args.no_extras:
        ne = min(n, 1 << 20)
        d_sk, d_nonce = torch.from_numpy(sk[:ne].view(np.int32)).to(dev), torch.from_numpy(nonce[:ne].view(np.int32)).to(dev)
        d_uo = torch.empty((ne, 8), dtype=torch.int32, device=dev)
        d_Ro, d_Ro2 = torch.empty((ne, 16), dtype=torch.int32, device=dev), torch.empty((ne, 16), dtype=torch.int32, device=dev)
        d_co = torch.empty((ne, 8), dtype=torch.int32, device=dev)

        def timed(fn, reps=3):
            fn(); fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                fn()
            b.record(stream)
            torch.cuda.synchronize()
            return world * ne / (max_over_ranks(a.elapsed_time(b) / reps) * 1e-3)

        extras["signs_per_s"] = timed(lambda: eng.call("sign", ne, fl, P(d_sk), P(d_msg), P(d_nonce), P(d_uo), P(d_Ro), P(d_co)))
        extras["sign_double_per_s"] = timed(lambda: eng.call("sign_double", ne, fl, P(d_sk), P(d_msg), P(d_nonce), P(d_uo), P(d_Ro), P(d_Ro2), P(d_co)))
        extras["verify_double_per_s"] = timed(lambda: eng.call("verify_double", ne, fl, P(d_pk), P(d_pk), P(d_u), P(d_R), P(d_R), P(d_msg), P(d_bm), None))
        extras["verify_vargen_per_s"] = timed(lambda: eng.call("verify_vargen", ne, fl, P(d_pk), P(d_pk), P(d_u), P(d_R), P(d_msg), P(d_bm), None))
        extras["verify_dual_pipe_per_s"] = timed(lambda: eng.call("verify", ne, fl | VERIFY_DUAL_PIPE, P(d_pk), P(d_u), P(d_R), P(d_msg), P(d_bm), None))
        extras["keygen_per_s"] = timed(lambda: eng.call("keygen", ne, fl, P(d_sk), P(d_Ro)))
        extras["batch"] = ne
        extras["note"] = "device-resident, CUDA events, per-op kernels of the same library; verdict content not checked here"

    # ---- CPU baseline beside it (rank 0, N=1 only) + verdict cross-check ---------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ref_cpu
        cores = os.cpu_count() or 1
        ns = args.ref_sample or 1024 * cores
        t0 = time.perf_counter()
        okc, _ = ref_cpu.verify(pk[:ns], u[:ns], R[:ns], msg[:ns])
        dt = time.perf_counter() - t0
        assert (okc == verdict[:ns]).all(), "GPU and CPU-restatement verdicts differ"
        cpu = {"value": ns / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {ns} tuples of the batch, {cores} pthreads; verdicts identical to the GPU's"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": f"single-key PublicKey::verify, batch 2^{args.log2_batch} per GPU, 10% corrupted signatures, "
                                   "affine (u,v) Montgomery inputs; sharded by tuple index, no collective",
                       "batch_per_gpu": n, "l2": "inputs (1 GiB) larger than L2; no flush needed"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "timing": "wall clock around the blocking host-buffer call (pinned host memory, chunked 2-stream pipeline)"},
            "gpu_launches": launches,
            "clocks": sampler.result(),
            "roofline": roof,
            "cpu_baseline": cpu,
            "other_ops": extras,
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
