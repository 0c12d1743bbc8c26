"""Robustness of the C ABI on the GPU: the full-size fallback of the half-size-scalar
path driven deterministically, checked input points, byte-level signers rejecting non-canonical scalars, pageable
vs pinned host buffers through the staging ring, stream switches, error returns that leave the context usable."""
import ctypes
import random

import numpy as np
import pytest

import schnorr_oracle as o
import vectors as V

pytestmark = pytest.mark.gpu
Q, R = o.Q, o.R


def test_half_size_fallback_forced_on_chosen_lanes(engine):
    """sb200_dbg_verify_ec with caller-chosen challenges: every third lane gets a c for which hgcd.cuh reports
    `ok = false`, so each warp takes the vote and runs verify_ec_core beside the fast path; keys carry torsion."""
    rnd = random.Random(41)
    n = 101
    hostile = V.hgcd_hostile_challenges(n // 3 + 1)
    tors = V.torsion_points()
    pk, u, Rr, cs, want = [], [], [], [], []
    for i in range(n):
        c = hostile[i // 3] if i % 3 == 0 else rnd.randrange(1 << 250)
        P = V.mul(o.G, rnd.randrange(R))
        if i % 5 == 1:
            P = o.pt_add(P, tors[i % len(tors)])  # key outside the prime-order subgroup
        ui = rnd.randrange(R)
        good = o.pt_add(V.mul(o.G, ui), V.mul(P, c))
        corrupt = i % 4 == 2
        pk.append(P); u.append(ui); cs.append(c)
        Rr.append(o.pt_add(good, o.G) if corrupt else good)
        want.append(not corrupt)
    got = engine.dbg_verify_ec(V.points(pk), V.scalars(u), V.points(Rr), V.scalars(cs))
    assert got.tolist() == want
    zs = [rnd.randrange(1, Q) for _ in range(n)]
    got = engine.dbg_verify_ec(V.points(pk, zs), V.scalars(u), V.points(Rr, zs[::-1]), V.scalars(cs), affine=False)
    assert got.tolist() == want


def test_points_check_and_check_points_flag(engine):
    rnd = random.Random(42)
    n = 64
    sk, nonce, msg = ([rnd.randrange(R) for _ in range(n)] for _ in range(3))
    msg = [m % Q for m in msg]
    pk = [V.mul(o.G, a) for a in sk]
    sig = [o.sign(a, b, m, mul=V.mul) for a, b, m in zip(sk, nonce, msg)]
    bad_pk = list(pk)
    off = set(range(3, n, 7))
    for i in off:
        bad_pk[i] = (pk[i][0], (pk[i][1] + 1) % Q)  # off the curve
        assert not o.on_curve(bad_pk[i])
    assert engine.points_check(V.points(bad_pk)).tolist() == [i not in off for i in range(n)]
    # projective: Z = 0 is rejected, any other Z accepted
    zs = [rnd.randrange(1, Q) for _ in range(n)]
    proj = V.points(pk, zs)
    proj[5, 16:24] = 0
    assert engine.points_check(proj, affine=False).tolist() == [i != 5 for i in range(n)]
    u, Rr = V.scalars([s[0] for s in sig]), V.points([s[1] for s in sig])
    ok_plain, _ = engine.verify(V.points(pk), u, Rr, V.fqs(msg))
    ok_chk, _ = engine.verify(V.points(bad_pk), u, Rr, V.fqs(msg), check_points=True)
    assert ok_plain.all() and ok_chk.tolist() == [i not in off for i in range(n)]
    bad_R = V.points([s[1] for s in sig]).copy()
    bad_R[9, 8:16] = V.mont(5)
    ok_chk, _ = engine.verify(V.points(pk), u, bad_R, V.fqs(msg), check_points=True)
    assert ok_chk.tolist() == [i != 9 for i in range(n)]
    # the other two schemes
    pkp = [V.mul(o.G_NUMS, a) for a in sk]
    sd = [o.sign_double(a, b, m, mul=V.mul) for a, b, m in zip(sk, nonce, msg)]
    bad_pkp = list(pkp)
    bad_pkp[11] = (1, 1)
    okd, _ = engine.verify_double(V.points(pk), V.points(bad_pkp), V.scalars([s[0] for s in sd]), V.points([s[1] for s in sd]),
                                  V.points([s[2] for s in sd]), V.fqs(msg), check_points=True)
    assert okd.tolist() == [i != 11 for i in range(n)]
    gens = [V.mul(o.G, rnd.randrange(1, R)) for _ in range(n)]
    sv = [o.sign_vargen(a, g, b, m, mul=V.mul) for a, g, b, m in zip(sk, gens, nonce, msg)]
    pkv = [V.mul(g, a) for g, a in zip(gens, sk)]
    bad_g = V.points(gens, zs)
    bad_g[13, 16:24] = 0  # Z = 0
    okv, _ = engine.verify_vargen(V.points(pkv, zs), bad_g, V.scalars([s[0] for s in sv]), V.points([s[1] for s in sv], zs),
                                  V.fqs(msg), affine=False, check_points=True)
    assert okv.tolist() == [i != 13 for i in range(n)]


def _b32(xs):
    return np.frombuffer(b"".join(int(x).to_bytes(32, "little") for x in xs), dtype=np.uint8).reshape(-1, 32)


def test_byte_level_signers_reject_non_canonical_scalars(engine):
    """SecretKey::from_bytes / JubJubScalar::from_bytes / BlsScalar::from_bytes -> Err(InvalidData): invalid bit set and
    an all-zero signature, in all three signers alike; the neighbours in the same warp are untouched"""
    rnd = random.Random(43)
    n = 40
    sk, nonce = [rnd.randrange(R) for _ in range(n)], [rnd.randrange(R) for _ in range(n)]
    msg = [rnd.randrange(Q) for _ in range(n)]
    bad = {3: "sk", 8: "nonce", 9: "nonce_huge", 17: "msg", 33: "sk_max"}
    sk[3] = R
    nonce[8] = R + 5
    nonce[9] = (1 << 256) - 1  # would overflow the window recoding if it were not replaced
    msg[17] = Q
    sk[33] = (1 << 256) - 1
    want_inv = [i in bad for i in range(n)]
    sig, inv = engine.sign_bytes(_b32(sk), _b32(msg), _b32(nonce), want_invalid=True)
    assert inv.tolist() == want_inv
    for i in range(n):
        if i in bad:
            assert not sig[i].any()
        else:
            u, Rp, _ = o.sign(sk[i], nonce[i], msg[i], mul=V.mul)
            assert bytes(sig[i]) == u.to_bytes(32, "little") + o.affine_to_bytes(Rp)
    sigd, invd = engine.sign_double_bytes(_b32(sk), _b32(msg), _b32(nonce), want_invalid=True)
    assert invd.tolist() == want_inv and all(not sigd[i].any() for i in bad)
    u, R1, R2, _ = o.sign_double(sk[0], nonce[0], msg[0], mul=V.mul)
    assert bytes(sigd[0]) == u.to_bytes(32, "little") + o.affine_to_bytes(R1) + o.affine_to_bytes(R2)
    gen = V.mul(o.G, 77)
    sk64 = np.concatenate([_b32(sk), np.tile(np.frombuffer(o.affine_to_bytes(gen), np.uint8), (n, 1))], axis=1).copy()
    sk64[21, 32:] = np.frombuffer((2).to_bytes(32, "little"), np.uint8)  # generator that does not decode
    sigv, okv = engine.sign_vargen_bytes(sk64, _b32(msg), _b32(nonce))
    assert okv.tolist() == [not w and i != 21 for i, w in enumerate(want_inv)]
    assert all(not sigv[i].any() for i in list(bad) + [21])
    u, Rp, _ = o.sign_vargen(sk[1], gen, nonce[1], msg[1], mul=V.mul)
    assert bytes(sigv[1]) == u.to_bytes(32, "little") + o.affine_to_bytes(Rp)


def test_pageable_and_pinned_host_buffers_agree(engine):
    """numpy (pageable) buffers go through the library's pinned staging ring, sb200_host_alloc buffers are copied
    directly; several pipeline chunks either way, inputs and outputs (sign writes 96 B per tuple back)"""
    from schnorr_b200 import POINTS_AFFINE, PinnedBuffer
    n = 3 * (1 << 16) + 77  # chunk = 2^16: 4 chunks, the last ragged
    rs = np.random.RandomState(5)
    def sc(bits):
        a = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
        a[:, 7] &= (1 << bits) - 1
        return a
    sk, nonce, msg = sc(27), sc(27), sc(30)
    pk = engine.keygen(sk)              # pageable in, pageable out
    u, Rr, c = engine.sign(sk, msg, nonce)
    bufs = [PinnedBuffer(s) for s in ((n, 8), (n, 8), (n, 8), (n, 8), (n, 16), (n, 8), (n, 16), ((n + 31) // 32,))]
    p_sk, p_msg, p_nonce, p_u, p_R, p_c, p_pk, p_bm = (b.array for b in bufs)
    p_sk[...] = sk; p_msg[...] = msg; p_nonce[...] = nonce
    P = lambda a: a.ctypes.data
    engine.call("sign", n, 0, P(p_sk), P(p_msg), P(p_nonce), P(p_u), P(p_R), P(p_c))
    engine.call("keygen", n, 0, P(p_sk), P(p_pk))
    assert (p_u == u).all() and (p_R == Rr).all() and (p_c == c).all() and (p_pk == pk).all()
    u2 = u.copy()
    u2[::9, 0] ^= 1
    p_u[...] = u2
    p_bm[...] = 0
    engine.call("verify", n, POINTS_AFFINE, P(p_pk), P(p_u), P(p_R), P(p_msg), P(p_bm), None)
    ok_pinned = np.unpackbits(p_bm.view(np.uint8), bitorder="little")[:n].astype(bool)
    ok_pageable, _ = engine.verify(pk, u2, Rr, msg, want_c=False)
    # mixed: pinned inputs, pageable verdict bitmap
    bm = np.zeros((n + 31) // 32 + 4, np.uint32)
    off = (-bm.ctypes.data // 4) % 4
    engine.call("verify", n, POINTS_AFFINE, P(p_pk), P(p_u), P(p_R), P(p_msg), bm[off:].ctypes.data, None)
    ok_mixed = np.unpackbits(bm[off:off + (n + 31) // 32].view(np.uint8), bitorder="little")[:n].astype(bool)
    want = (np.arange(n) % 9) != 0
    assert (ok_pinned == want).all() and (ok_pageable == want).all() and (ok_mixed == want).all()
    for b in bufs:
        b.close()


def test_device_pointer_calls_on_two_streams_are_ordered(engine):
    """SB200_DEVICE_PTRS calls share the context's scratch rows; after sb200_set_stream the new stream waits for the
    work enqueued on the old one, so back-to-back calls on different streams cannot overwrite each other's challenges"""
    import torch
    from schnorr_b200 import DEVICE_PTRS, POINTS_AFFINE
    dev = torch.device("cuda:0")
    n = 1 << 15
    rs = np.random.RandomState(6)
    def sc(bits):
        a = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
        a[:, 7] &= (1 << bits) - 1
        return a
    batches = []
    for k in range(2):
        sk, nonce, msg = sc(27), sc(27), sc(30)
        pk = engine.keygen(sk)
        u, Rr, _ = engine.sign(sk, msg, nonce)
        u[k::5, 0] ^= 1
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).to(dev)
        batches.append((t(pk), t(u), t(Rr), t(msg), torch.zeros((n + 31) // 32, dtype=torch.int32, device=dev)))
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    P = lambda x: x.data_ptr()
    for rep in range(3):
        for k, (pk, u, Rr, msg, bm) in enumerate(batches):
            engine.set_stream(streams[k].cuda_stream)
            engine.call("verify", n, POINTS_AFFINE | DEVICE_PTRS, P(pk), P(u), P(Rr), P(msg), P(bm), None)
    torch.cuda.synchronize()
    engine.set_stream(0)
    for k, b in enumerate(batches):
        ok = np.unpackbits(b[4].cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
        assert (ok == ((np.arange(n) - k) % 5 != 0)).all()


def test_errors_are_codes_and_leave_the_context_usable(engine):
    from schnorr_b200 import SchnorrB200Error, _lib
    a = np.zeros(8 * 64 + 4, np.uint32)
    mis = a[1:] if a.ctypes.data % 16 == 0 else a  # a pointer that is NOT 16-byte aligned
    while mis.ctypes.data % 16 == 0:
        mis = mis[1:]
    good = _lib.aligned_empty((64, 16))
    rc = engine._lib.sb200_keygen(engine._h, 64, 0, ctypes.c_void_p(mis.ctypes.data), ctypes.c_void_p(good.ctypes.data))
    assert rc == _lib.ERR_ARG
    rc = engine._lib.sb200_keygen(engine._h, -1, 0, ctypes.c_void_p(good.ctypes.data), ctypes.c_void_p(good.ctypes.data))
    assert rc == _lib.ERR_ARG
    rc = engine._lib.sb200_verify(engine._h, 64, 0x40, None, None, None, None, None, None)  # unknown flag / null verdicts
    assert rc == _lib.ERR_ARG
    with pytest.raises(SchnorrB200Error):
        engine.call("keygen", 64, 0x80, good.ctypes.data, good.ctypes.data)
    pk = engine.keygen(V.scalars([5, 6, 7]))
    assert V.points_out(pk) == [V.mul(o.G, k) for k in (5, 6, 7)]


def test_oblivious_signing_and_keygen_match_default(engine):
    """SB200_SIGN_OBLIVIOUS (4-bit combs in shared memory read by masked scan; scanned window tables): bit-identical
    signatures / keys to the default path and to the oracle, for all three schemes"""
    rnd = random.Random(44)
    n = 300  # not a multiple of the CTA's tuple count: ragged tail
    sk = [0, 1, R - 1] + [rnd.randrange(R) for _ in range(n - 3)]
    nonce = [1, R - 1, 0] + [rnd.randrange(R) for _ in range(n - 3)]
    msg = [rnd.randrange(Q) for _ in range(n)]
    S, M, N = V.scalars(sk), V.fqs(msg), V.scalars(nonce)
    a, b = engine.sign(S, M, N), engine.sign(S, M, N, oblivious=True)
    assert all((x == y).all() for x, y in zip(a, b))
    for i in (0, 1, 2, 17):
        u, Rp, c = o.sign(sk[i], nonce[i], msg[i], mul=V.mul)
        assert (V.to_int(b[0][i]), V.points_out(b[1][i])[0], V.to_int(b[2][i])) == (u, Rp, c)
    a, b = engine.sign_double(S, M, N), engine.sign_double(S, M, N, oblivious=True)
    assert all((x == y).all() for x, y in zip(a, b))
    assert (engine.keygen(S) == engine.keygen(S, oblivious=True)).all()
    a, b = engine.keygen_double(S), engine.keygen_double(S, oblivious=True)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
    assert V.points_out(b[1][:3]) == [V.mul(o.G_NUMS, k) for k in sk[:3]]
    gens = [V.rand_curve_point(rnd) for _ in range(n)]  # whole curve
    zs = [rnd.randrange(1, Q) for _ in range(n)]
    G = V.points(gens, zs)
    a, b = engine.sign_vargen(S, G, M, N, affine=False), engine.sign_vargen(S, G, M, N, affine=False, oblivious=True)
    assert all((x == y).all() for x, y in zip(a, b))
    assert (engine.keygen_vargen(S, G, affine=False) == engine.keygen_vargen(S, G, affine=False, oblivious=True)).all()


def test_vargen_short_scalars_and_their_fallback(engine):
    """csrc/lat3.cuh on the GPU: the three-table short-scalar form of PublicKeyVarGen::verify against the oracle, with
    lanes that force the full-size fallback (u = r - 1: every short lattice vector has an even b) mixed into every
    warp -- valid and corrupted, generators and keys with torsion components"""
    rnd = random.Random(45)
    n = 130
    tors = V.torsion_points()
    pk, gen, u, Rr, msg, want = [], [], [], [], [], []
    for i in range(n):
        g = V.mul(o.G, rnd.randrange(1, R))
        m = rnd.randrange(Q)
        if i % 4 == 0:  # forced fallback, valid by construction: PK = (k - u) / c * g for R = k g
            ui, k = R - 1, rnd.randrange(1, R)
            Rp = V.mul(g, k)
            c = o.challenge_hash(Rp, m)
            P = V.mul(g, (k - ui) * pow(c, -1, R) % R)
        else:
            sk, nonce = rnd.randrange(1, R), rnd.randrange(R)
            ui, Rp, c = o.sign_vargen(sk, g, nonce, m, mul=V.mul)
            P = V.mul(g, sk)
        valid = True
        if i % 3 == 1:
            Rp = o.pt_add(Rp, g); valid = False      # wrong nonce point
        if i % 5 == 2:
            P = o.pt_add(P, tors[i % len(tors)])      # torsion component on the key: verdict from the oracle
            valid = None
        if i % 7 == 3:
            g = o.pt_add(g, tors[(i + 1) % len(tors)])
            valid = None
        if valid is None:
            valid = o.verify_vargen(P, g, ui, Rp, m, mul=V.mul)
        pk.append(P); gen.append(g); u.append(ui); Rr.append(Rp); msg.append(m); want.append(bool(valid))
    ok, _ = engine.verify_vargen(V.points(pk), V.points(gen), V.scalars(u), V.points(Rr), V.fqs(msg))
    assert ok.tolist() == want and any(want) and not all(want)
    zs = [rnd.randrange(1, Q) for _ in range(n)]
    ok, _ = engine.verify_vargen(V.points(pk, zs), V.points(gen, zs[::-1]), V.scalars(u), V.points(Rr, zs), V.fqs(msg), affine=False)
    assert ok.tolist() == want


@pytest.mark.gpu
def test_lattice3_on_the_device(engine):
    """csrc/lat3.cuh as compiled for the GPU (sb200_dbg_lattice3): every reported short vector (b, a, d) satisfies
    a = b c, d = b u (mod 8r) with b odd and all three below the 174-bit window budget -- edge, structured and random
    (c, u), several warps, lanes with `ok` = false (u = r - 1) mixed in; random inputs must essentially always fit"""
    rnd = random.Random(97)
    N = 8 * R
    edge = [0, 1, 2, R - 1, (1 << 250) - 1, N // 3 % (1 << 250), 1 << 128, (1 << 170) + 1]
    cases = [(c, u) for c in edge for u in (0, 1, R - 1, rnd.randrange(R))]
    n_edge = len(cases)
    cases += [(rnd.randrange(1 << 250), rnd.randrange(R)) for _ in range(1000)]
    cases += [(c, rnd.randrange(R)) for c in V.hgcd_hostile_challenges(4)]
    out = engine.dbg_lattice3(V.scalars([c for c, _ in cases]), V.scalars([u for _, u in cases]))
    n_ok = 0
    for k, ((c, u), row) in enumerate(zip(cases, out)):
        a, b, d = V.to_int(row[:8]), V.to_int(row[8:16]), V.to_int(row[16:24])
        a, b, d = (-a if row[24] else a), (-b if row[25] else b), (-d if row[26] else d)
        if not row[27]:
            continue
        n_ok += n_edge <= k < n_edge + 1000
        assert b % 2 == 1, (k, hex(c), hex(u))
        assert (a - b * c) % N == 0 and (d - b * u) % N == 0, (k, hex(c), hex(u), hex(a), hex(b), hex(d))
        assert max(abs(a), abs(b), abs(d)) < (1 << 174)
    assert n_ok >= 998


@pytest.mark.parametrize("mode", ["always", "never"])
def test_curve_kernel_forms_agree(mode, monkeypatch):
    """the two forms of the curve kernels -- persistent with global thread-major window tables staged through shared memory
    (what large batches run) and one-shot with local-memory tables (small batches) -- forced by SB200_CURVE_PERSISTENT on a
    second context: single-key, double-key and variable-generator verdicts against the oracle on a ragged batch with
    corrupted, torsion-shifted and small-order inputs, affine and projective"""
    from schnorr_b200 import Engine
    monkeypatch.setenv("SB200_CURVE_PERSISTENT", mode)
    eng = Engine([0])
    try:
        rnd = random.Random(46)
        n = 333
        tors = V.torsion_points()
        sk = [rnd.randrange(1, R) for _ in range(n)]
        nonce = [rnd.randrange(R) for _ in range(n)]
        msg = [rnd.randrange(Q) for _ in range(n)]
        # single-key
        pk = [V.mul(o.G, s) for s in sk]
        sig = [o.sign(s, k, m, mul=V.mul) for s, k, m in zip(sk, nonce, msg)]
        u, Rr = [x[0] for x in sig], [x[1] for x in sig]
        for i in range(n):
            if i % 3 == 1:
                u[i] = (u[i] + 1) % R
            if i % 5 == 2:
                pk[i] = o.pt_add(pk[i], tors[i % len(tors)])
            if i % 7 == 3:
                Rr[i] = o.pt_add(Rr[i], tors[(i + 2) % len(tors)])
        want = [o.verify(pk[i], u[i], Rr[i], msg[i], mul=V.mul) for i in range(n)]
        ok, _ = eng.verify(V.points(pk), V.scalars(u), V.points(Rr), V.fqs(msg))
        assert ok.tolist() == want and any(want) and not all(want)
        zs = [rnd.randrange(1, Q) for _ in range(n)]
        ok, _ = eng.verify(V.points(pk, zs), V.scalars(u), V.points(Rr, zs[::-1]), V.fqs(msg), affine=False)
        assert ok.tolist() == want
        # double-key
        m2 = 97
        pkd = [(V.mul(o.G, s), V.mul(o.G_NUMS, s)) for s in sk[:m2]]
        sd = [o.sign_double(s, k, m, mul=V.mul) for s, k, m in zip(sk[:m2], nonce[:m2], msg[:m2])]
        ud, R1, R2 = [x[0] for x in sd], [x[1] for x in sd], [x[2] for x in sd]
        for i in range(m2):
            if i % 4 == 1:
                R2[i] = o.pt_add(R2[i], o.G)
            if i % 6 == 2:
                ud[i] = (ud[i] + 5) % R
        wantd = [o.verify_double(pkd[i][0], pkd[i][1], ud[i], R1[i], R2[i], msg[i], mul=V.mul) for i in range(m2)]
        ok, _ = eng.verify_double(V.points([p[0] for p in pkd]), V.points([p[1] for p in pkd]), V.scalars(ud), V.points(R1), V.points(R2), V.fqs(msg[:m2]))
        assert ok.tolist() == wantd and any(wantd) and not all(wantd)
        # variable generator (with forced full-size fallback lanes: u = r - 1)
        gens = [V.mul(o.G, rnd.randrange(1, R)) for _ in range(m2)]
        pkv, uv, Rv = [], [], []
        for i in range(m2):
            g = gens[i]
            if i % 4 == 0:
                ui, k = R - 1, rnd.randrange(1, R)
                Rp = V.mul(g, k)
                c = o.challenge_hash(Rp, msg[i])
                P = V.mul(g, (k - ui) * pow(c, -1, R) % R)
            else:
                ui, Rp, c = o.sign_vargen(sk[i], g, nonce[i], msg[i], mul=V.mul)
                P = V.mul(g, sk[i])
            if i % 3 == 1:
                Rp = o.pt_add(Rp, g)
            pkv.append(P); uv.append(ui); Rv.append(Rp)
        wantv = [o.verify_vargen(pkv[i], gens[i], uv[i], Rv[i], msg[i], mul=V.mul) for i in range(m2)]
        ok, _ = eng.verify_vargen(V.points(pkv), V.points(gens), V.scalars(uv), V.points(Rv), V.fqs(msg[:m2]))
        assert ok.tolist() == wantv and any(wantv) and not all(wantv)
    finally:
        eng.close()
