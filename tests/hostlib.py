"""ctypes driver for build/host_arith.so (CPU build of the CUDA arithmetic headers; test scaffolding)."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Q = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
RADIX = 1 << 256


def build():
    so = os.path.join(ROOT, "build", "host_arith.so")
    src = os.path.join(ROOT, "tests", "host_arith.cpp")
    hdrs = [os.path.join(ROOT, "schnorr_b200", "csrc", f) for f in (
        "fq.cuh", "experimental/fd.cuh", "ed.cuh", "hades.cuh", "experimental/hades_fd.cuh", "core.cuh", "wire.cuh", "hgcd.cuh", "lat3.cuh", "inv.cuh",
        "params_host.cuh", "constants_gen.cuh", "experimental/constants_fd_gen.cuh")]
    newest = max(os.path.getmtime(p) for p in [src] + hdrs)
    if not os.path.exists(so) or os.path.getmtime(so) < newest:
        os.makedirs(os.path.dirname(so), exist_ok=True)
        # SB_EXPERIMENTAL_FD=1: the host tests also cover the compiled-out FP64-pipe arithmetic
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-DSB_EXPERIMENTAL_FD=1", "-include",
                               os.path.join(ROOT, "tests", "host_shim.h"), "-o", so, src])
    lib = ctypes.CDLL(so)
    setup_hades(lib)
    return lib


def setup_hades(lib, rule=None):
    """Hand the oracle's round constants / MDS (current rule, or `rule`) to the host build, which derives its sparse
    tables from them the way sb200_init_ex does."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import schnorr_oracle as o
    if rule is not None and rule != o.ARK_RULE:
        o.set_ark_rule(rule)
    rc = np.stack([mont(c) for c in o.ROUND_CONSTANTS[:335]])
    mds = np.stack([mont(o.MDS[i][j]) for i in range(5) for j in range(5)])
    assert lib.h_setup_hades(ptr(rc), ptr(mds)) == 0


def limbs(x):
    return np.frombuffer(int(x).to_bytes(32, "little"), dtype=np.uint32).copy()


def to_int(a):
    return int.from_bytes(np.ascontiguousarray(a, dtype=np.uint32).tobytes(), "little")


def mont(x):
    return limbs(x * RADIX % Q)


def unmont(a):
    return to_int(a) * pow(RADIX, -1, Q) % Q


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def pt_mont(p, z=None):
    """affine (u, v) [or projective with z] -> contiguous Montgomery limbs"""
    if z is None:
        return np.concatenate([mont(p[0]), mont(p[1])])
    return np.concatenate([mont(p[0] * z % Q), mont(p[1] * z % Q), mont(z)])


def comb_tables(lib):
    """(G table, G' table) as built by the host build of comb_build_entry; cached on disk (slow emulation)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import schnorr_oracle as o
    cache = os.path.join(ROOT, "build", "host_comb_tables.npy")
    hdr = os.path.join(ROOT, "schnorr_b200", "csrc", "ed.cuh")
    if os.path.exists(cache) and os.path.getmtime(cache) > os.path.getmtime(hdr):
        t = np.load(cache)
        return t[0].copy(), t[1].copy()
    tabs = []
    for B in (o.G, o.G_NUMS):
        t = np.zeros(32 * 129 * 24, np.uint32)
        lib.h_comb_build(ptr(mont(B[0])), ptr(mont(B[1])), ptr(t))
        tabs.append(t)
    np.save(cache, np.stack(tabs))
    return tabs[0], tabs[1]
