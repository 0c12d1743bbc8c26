"""Scheme parameters as inputs of context creation (include/schnorr_b200.h `sb200_params`).

CPU part (no GPU call): the library's recipe-derived defaults equal the oracle's constants under both recalled
round-constant rules; the C++ derivation of the sparse Hades tables equals its Python twin (tools/gen_constants.py) bit
for bit; malformed parameters are rejected.  GPU part: contexts created with either rule (and with arbitrary
caller-supplied tables) reproduce the oracle's permutation / challenges; one device holds one parameter set."""
import ctypes
import os
import random
import sys

import numpy as np
import pytest

import schnorr_oracle as o
import vectors as V

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
Q, R = o.Q, o.R
MONT = 1 << 256


def _ints(a):
    return [V.to_int(r) for r in np.asarray(a).reshape(-1, 8)]


def _copy(p):
    from schnorr_b200 import Params
    q = Params()
    ctypes.memmove(ctypes.byref(q), ctypes.byref(p), ctypes.sizeof(Params))
    return q


@pytest.mark.parametrize("rule", ["cumsum", "plain"])
def test_default_params_match_oracle(rule):
    from schnorr_b200 import _lib
    p = _lib.default_params(_lib.ark_rule(rule))
    assert p.struct_size == ctypes.sizeof(_lib.Params) == 8 + 32 * (4 + 335 + 25)
    mont = lambda x: x * MONT % Q
    assert _ints(p.view("generator")) == [mont(o.G[0]), mont(o.G[1])]
    assert _ints(p.view("generator_nums")) == [mont(o.G_NUMS[0]), mont(o.G_NUMS[1])]
    assert _ints(p.view("round_constants")) == [mont(c) for c in o._round_constants(rule)[:335]]
    assert _ints(p.view("mds")) == [mont(o.MDS[i][j]) for i in range(5) for j in range(5)]


def test_default_rule_follows_environment(monkeypatch):
    from schnorr_b200 import _lib
    monkeypatch.setenv("SB200_ARK", "plain")
    assert _ints(_lib.default_params().view("round_constants"))[0] == o._round_constants("plain")[0] * MONT % Q
    monkeypatch.setenv("SB200_ARK", "cumsum")
    assert _ints(_lib.default_params().view("round_constants"))[0] == o._round_constants("cumsum")[0] * MONT % Q
    monkeypatch.setenv("SB200_ARK", "nonsense")
    with pytest.raises(_lib.SchnorrB200Error):
        _lib.default_params()


@pytest.mark.parametrize("rule", ["cumsum", "plain"])
def test_derived_tables_equal_python_twin(rule):
    """csrc/params_host.cuh derive_hades_tables == tools/gen_constants.py build_sparse_partial (product side, both)"""
    import gen_constants as G
    from schnorr_b200 import _lib
    rc, tabs = _lib.params_check(_lib.default_params(_lib.ark_rule(rule)))
    assert rc == 0
    want = G.hades_tables(rule)
    G.set_ark(G.DEFAULT_ARK)
    for k in ("rc", "mds", "pre", "sparse", "post"):
        assert _ints(tabs[k]) == want[k], k


def test_arbitrary_tables_are_accepted_and_bad_ones_rejected():
    """nothing about the recalled constants is special-cased: any canonical round constants + an MDS matrix derive"""
    from schnorr_b200 import _lib
    rnd = random.Random(11)
    p = _lib.default_params(_lib.ARK_PLAIN)
    p.view("round_constants")[...] = np.stack([V.limbs(rnd.randrange(Q)) for _ in range(335)])
    xs, ys = [rnd.randrange(Q) for _ in range(5)], [rnd.randrange(Q) for _ in range(5)]
    p.view("mds")[...] = np.stack([V.mont(pow(xs[i] + ys[j], -1, Q)) for i in range(5) for j in range(5)])  # another Cauchy matrix
    rc, tabs = _lib.params_check(p)
    assert rc == 0 and _ints(tabs["rc"]) == _ints(p.view("round_constants"))

    def rejected(mutate, code=_lib.ERR_PARAMS):
        q = _copy(_lib.default_params(_lib.ARK_CUMSUM))
        mutate(q)
        assert _lib.params_check(q)[0] == code

    rejected(lambda q: q.view("round_constants").__setitem__((5,), V.limbs(Q)))            # non-canonical field element
    rejected(lambda q: q.view("generator").__setitem__((0,), V.mont(o.G[0] + 1)))          # off the curve
    rejected(lambda q: q.view("generator_nums").__setitem__(slice(None), np.stack([V.mont(0), V.mont(1)])))   # identity
    t8 = next(t for t in V.torsion_points() if o.pt_mul_fast(t, 4) != o.IDENTITY)
    rejected(lambda q: q.view("generator").__setitem__(slice(None), np.stack([V.mont(t8[0]), V.mont(t8[1])])))  # order 8
    mixed = o.pt_add(o.G, t8)
    rejected(lambda q: q.view("generator").__setitem__(slice(None), np.stack([V.mont(mixed[0]), V.mont(mixed[1])])))  # order 8r
    rejected(lambda q: q.view("mds").__setitem__((6,), q.view("mds")[1].copy()) or q.view("mds").__setitem__((5,), q.view("mds")[0].copy())
             or [q.view("mds").__setitem__((5 + j,), q.view("mds")[j].copy()) for j in range(5)])  # two equal rows: singular block
    rejected(lambda q: setattr(q, "struct_size", 16), _lib.ERR_ARG)


# ----------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_contexts_under_both_rules(ark, ark_engine):
    """a context created with either rule reproduces the oracle under that rule: permutation (dense and sparse),
    challenges, signatures, verdicts"""
    e, rnd = ark_engine, random.Random(21)
    assert _ints(e.dbg_hades_tables()["rc"]) == [c * MONT % Q for c in o.ROUND_CONSTANTS[:335]]
    states = [[rnd.randrange(Q) for _ in range(5)] for _ in range(33)] + [[0] * 5]
    arr = np.stack([np.concatenate([V.mont(x) for x in s]) for s in states])
    want = [o.hades_perm(s) for s in states]
    for dense in (False, True):
        out = e.dbg_hades(arr, dense=dense)
        assert [[V.unmont(r[8 * k:8 * k + 8]) for k in range(5)] for r in out] == want
    n = 70
    sk, nonce, msg = ([rnd.randrange(R) for _ in range(n)], [rnd.randrange(R) for _ in range(n)], [rnd.randrange(Q) for _ in range(n)])
    u, Rr, c = e.sign(V.scalars(sk), V.fqs(msg), V.scalars(nonce))
    exp = [o.sign(a, b, m, mul=V.mul) for a, b, m in zip(sk, nonce, msg)]
    assert V.ints_out(u) == [x[0] for x in exp] and V.points_out(Rr) == [x[1] for x in exp] and V.ints_out(c) == [x[2] for x in exp]
    u2 = u.copy()
    u2[::3, 0] ^= 1
    ok, c2 = e.verify(e.keygen(V.scalars(sk)), u2, Rr, V.fqs(msg))
    assert ok.tolist() == [i % 3 != 0 for i in range(n)] and V.ints_out(c2) == [x[2] for x in exp]
    ud, R1, R2, cd = e.sign_double(V.scalars(sk), V.fqs(msg), V.scalars(nonce))
    expd = [o.sign_double(a, b, m, mul=V.mul) for a, b, m in zip(sk, nonce, msg)]
    assert V.ints_out(ud) == [x[0] for x in expd] and V.ints_out(cd) == [x[3] for x in expd]


@pytest.mark.gpu
def test_caller_supplied_tables_and_one_parameter_set_per_device(engine):
    """(1) a second context with DIFFERENT parameters on the same device is refused (constant memory is shared);
    (2) after the first is closed, a context with caller-chosen (non-default) round constants and generator runs the
    permutation / a signature the oracle computes with the same values"""
    import conftest
    from schnorr_b200 import Engine, SchnorrB200Error, _lib
    other = "plain" if o.ARK_RULE == "cumsum" else "cumsum"
    with pytest.raises(SchnorrB200Error) as ei:
        Engine([0], ark=other)
    assert ei.value.code == _lib.ERR_BUSY
    same = Engine([0], ark=o.ARK_RULE)  # same parameters: fine
    same.close()
    conftest.engine_for(None) if False else None
    conftest._live["engine"].close()
    conftest._live["engine"] = None
    rnd = random.Random(31)
    p = _lib.default_params(_lib.ARK_PLAIN)
    rcs = [rnd.randrange(Q) for _ in range(335)]
    p.view("round_constants")[...] = np.stack([V.mont(x) for x in rcs])
    g2 = V.mul(o.G, 0xABCDEF)
    p.view("generator")[...] = np.stack([V.mont(g2[0]), V.mont(g2[1])])
    saved_rc, saved_g = list(o.ROUND_CONSTANTS), o.G
    e = Engine([0], params=p)
    try:
        o.ROUND_CONSTANTS[:335] = rcs
        o.G = g2
        sk, nonce, msg = rnd.randrange(R), rnd.randrange(R), rnd.randrange(Q)
        u, Rr, c = e.sign(V.scalars([sk]), V.fqs([msg]), V.scalars([nonce]))
        want = o.sign(sk, nonce, msg, mul=V.mul)
        assert (V.ints_out(u)[0], V.points_out(Rr)[0], V.ints_out(c)[0]) == want
        ok, _ = e.verify(e.keygen(V.scalars([sk])), u, Rr, V.fqs([msg]))
        assert ok.all() and V.points_out(e.keygen(V.scalars([sk])))[0] == V.mul(g2, sk)
    finally:
        o.ROUND_CONSTANTS[:] = saved_rc
        o.G = saved_g
        e.close()
