"""ctypes driver for oracle/ref_cpu.c (multi-threaded CPU restatement of the reference's algorithm).

TEST INFRASTRUCTURE ONLY -- see the header of ref_cpu.c.  Arrays use the same layout as the GPU
library's ABI (uint32 limbs; field elements in Montgomery form, scalars canonical).
"""
import ctypes
import os
import subprocess

import numpy as np

import schnorr_oracle as o

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "ref_cpu.c")
SO = os.path.join(_HERE, "_build", "libref_cpu.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-fPIC", "-shared", "-pthread", "-o", SO, SRC])
    return SO


def _mont(x):
    return x * o.MONT_R % o.Q


def _w(x):
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


_rule = None


def lib():
    global _lib, _rule
    if _lib is None or _rule != o.ARK_RULE:  # (re)upload the constants: the oracle's round-constant rule may have changed
        build()
        L = _lib or ctypes.CDLL(SO)
        _rule = o.ARK_RULE
        blob = []
        blob += _w(o.Q) + _w(o.R)
        blob += [(-pow(o.Q, -1, 1 << 64)) % (1 << 64), (-pow(o.R, -1, 1 << 64)) % (1 << 64)]
        blob += _w(o.MONT_R % o.Q) + _w(o.MONT_R ** 2 % o.Q) + _w(o.MONT_R ** 2 % o.R)
        blob += _w(_mont(2 * o.D % o.Q))
        for p in (o.G, o.G_NUMS):
            blob += _w(_mont(p[0])) + _w(_mont(p[1]))
        for c in o.ROUND_CONSTANTS[:335]:
            blob += _w(_mont(c))
        for i in range(5):
            for j in range(5):
                blob += _w(_mont(o.MDS[i][j]))
        arr = np.array(blob, dtype=np.uint64)
        L.ref_init(arr.ctypes.data_as(ctypes.c_void_p))
        _lib = L
    return _lib


def _p(a):
    return None if a is None else np.ascontiguousarray(a).ctypes.data_as(ctypes.c_void_p)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def default_threads():
    return os.cpu_count() or 1


def verify(pk, u, R, m, affine=True, threads=None):
    pk, u, R, m = _c(pk), _c(u), _c(R), _c(m)
    n = u.size // 8
    v, c = np.zeros(n, np.uint8), np.zeros((n, 8), np.uint32)
    lib().ref_verify(ctypes.c_int64(n), int(affine), _p(pk), _p(u), _p(R), _p(m), _p(v), _p(c), threads or default_threads())
    return v.astype(bool), c


def verify_double(pk, pkp, u, R, Rp, m, affine=True, threads=None):
    pk, pkp, u, R, Rp, m = _c(pk), _c(pkp), _c(u), _c(R), _c(Rp), _c(m)
    n = u.size // 8
    v, c = np.zeros(n, np.uint8), np.zeros((n, 8), np.uint32)
    lib().ref_verify_double(ctypes.c_int64(n), int(affine), _p(pk), _p(pkp), _p(u), _p(R), _p(Rp), _p(m), _p(v), _p(c),
                            threads or default_threads())
    return v.astype(bool), c


def verify_vargen(pk, gen, u, R, m, affine=True, threads=None):
    pk, gen, u, R, m = _c(pk), _c(gen), _c(u), _c(R), _c(m)
    n = u.size // 8
    v, c = np.zeros(n, np.uint8), np.zeros((n, 8), np.uint32)
    lib().ref_verify_vargen(ctypes.c_int64(n), int(affine), _p(pk), _p(gen), _p(u), _p(R), _p(m), _p(v), _p(c),
                            threads or default_threads())
    return v.astype(bool), c


def sign(sk, m, nonce, threads=None):
    sk, m, nonce = _c(sk), _c(m), _c(nonce)
    n = sk.size // 8
    u, R, c = np.zeros((n, 8), np.uint32), np.zeros((n, 16), np.uint32), np.zeros((n, 8), np.uint32)
    lib().ref_sign(ctypes.c_int64(n), _p(sk), _p(m), _p(nonce), _p(u), _p(R), _p(c), threads or default_threads())
    return u, R, c


def sign_double(sk, m, nonce, threads=None):
    sk, m, nonce = _c(sk), _c(m), _c(nonce)
    n = sk.size // 8
    u, R, Rp, c = (np.zeros((n, 8), np.uint32), np.zeros((n, 16), np.uint32), np.zeros((n, 16), np.uint32),
                   np.zeros((n, 8), np.uint32))
    lib().ref_sign_double(ctypes.c_int64(n), _p(sk), _p(m), _p(nonce), _p(u), _p(R), _p(Rp), _p(c), threads or default_threads())
    return u, R, Rp, c


def sign_vargen(sk, gen, m, nonce, affine=True, threads=None):
    sk, gen, m, nonce = _c(sk), _c(gen), _c(m), _c(nonce)
    n = sk.size // 8
    u, R, c = np.zeros((n, 8), np.uint32), np.zeros((n, 16), np.uint32), np.zeros((n, 8), np.uint32)
    lib().ref_sign_vargen(ctypes.c_int64(n), int(affine), _p(sk), _p(gen), _p(m), _p(nonce), _p(u), _p(R), _p(c),
                          threads or default_threads())
    return u, R, c


def keygen(sk, gen=None, double=False, affine=True, threads=None):
    sk = _c(sk)
    gen = None if gen is None else _c(gen)
    n = sk.size // 8
    pk = np.zeros((n, 16), np.uint32)
    pkp = np.zeros((n, 16), np.uint32) if double else None
    lib().ref_keygen(ctypes.c_int64(n), int(affine), _p(sk), _p(gen), _p(pk), _p(pkp), threads or default_threads())
    return (pk, pkp) if double else pk


def scalar_mul(points, k, affine=True, threads=None):
    points, k = _c(points), _c(k)
    n = k.size // 8
    out = np.zeros((n, 16), np.uint32)
    lib().ref_scalar_mul(ctypes.c_int64(n), int(affine), _p(points), _p(k), _p(out), threads or default_threads())
    return out
