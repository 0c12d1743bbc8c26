"""GPU parity: every CUDA building block and every batch entry point against the oracle, bit-exact."""
import random

import numpy as np
import pytest

import schnorr_oracle as o
import vectors as V

pytestmark = pytest.mark.gpu
Q, R = o.Q, o.R


def edge_fq():
    return [0, 1, 2, Q - 1, Q - 2, (1 << 32) - 1, 1 << 32, (1 << 64) - 1, (1 << 255) % Q, Q >> 1, (Q >> 1) + 1,
            0xFFFFFFFF00000000FFFFFFFF, V.RADIX % Q, (V.RADIX * V.RADIX) % Q]


def test_fq_ops(engine):
    rnd = random.Random(11)
    xs = edge_fq() + [rnd.randrange(Q) for _ in range(500)]
    a = [x for x in xs for _ in range(len(edge_fq()) + 8)]
    b = [y for _ in xs for y in (edge_fq() + [rnd.randrange(Q) for _ in range(8)])]
    A, B = V.scalars(a), V.scalars(b)
    rinv = V.RINV
    assert V.ints_out(engine.dbg_fq(0, A, B)) == [x * y * rinv % Q for x, y in zip(a, b)]
    assert V.ints_out(engine.dbg_fq(1, A, B)) == [(x + y) % Q for x, y in zip(a, b)]
    assert V.ints_out(engine.dbg_fq(2, A, B)) == [(x - y) % Q for x, y in zip(a, b)]
    assert V.ints_out(engine.dbg_fq(4, A)) == [x * x * rinv % Q for x in a]
    assert V.ints_out(engine.dbg_fq(5, A)) == [x * V.RADIX % Q for x in a]
    assert V.ints_out(engine.dbg_fq(6, A)) == [x * rinv % Q for x in a]


def test_fq_inv(engine):
    rnd = random.Random(12)
    xs = [x for x in edge_fq() if x] + [rnd.randrange(1, Q) for _ in range(200)]
    got = V.ints_out(engine.dbg_fq(3, V.fqs(xs)))
    assert got == [pow(x, -1, Q) * V.RADIX % Q for x in xs]
    # 0 has no inverse: Fermat gives 0 (the reference would panic on Z = 0; documented precondition)
    assert V.ints_out(engine.dbg_fq(3, V.fqs([0]))) == [0]


def test_fq_inv_euclid(engine):
    """csrc/inv.cuh on the device (op 7): the two-level Euclidean inversion against the big-int inverse -- random values, edge
    values, and values whose Montgomery form is a SMALL integer (first quotient ~2^220: those lanes give up and are inverted by
    Fermat after a warp vote) mixed into the same warps; 0 -> 0"""
    rnd = random.Random(14)
    rinv = pow(V.RADIX, -1, Q)
    small = [k * rinv % Q for k in (1, 2, 3, 0xFFFFFFFF, 1 << 32, (1 << 64) + 5, (1 << 100) - 1, (1 << 213) + 12345)]
    xs = [x for x in edge_fq() if x] + small + [rnd.randrange(1, Q) for _ in range(500)]
    xs += [rnd.choice(small) if i % 5 == 0 else rnd.randrange(1, Q) for i in range(300)]
    got = V.ints_out(engine.dbg_fq(7, V.fqs(xs)))
    assert got == [pow(x, -1, Q) * V.RADIX % Q for x in xs]
    assert V.ints_out(engine.dbg_fq(7, V.fqs([0, 5, 0]))) == [0, pow(5, -1, Q) * V.RADIX % Q, 0]


def test_fr_mul(engine):
    rnd = random.Random(13)
    xs = [0, 1, R - 1, (1 << 250) - 1, 1 << 251] + [rnd.randrange(R) for _ in range(300)]
    ys = [rnd.choice(xs) for _ in xs]
    assert V.ints_out(engine.dbg_fr_mul(V.scalars(xs), V.scalars(ys))) == [x * y % R for x, y in zip(xs, ys)]


@pytest.mark.parametrize("dense", [True, False], ids=["dense", "sparse"])
def test_hades(engine, dense):
    rnd = random.Random(14)
    states = [[0] * 5, [1] * 5, [Q - 1] * 5, [0, 1, 2, 3, 4]] + [[rnd.randrange(Q) for _ in range(5)] for _ in range(60)]
    buf = np.stack([np.concatenate([V.mont(x) for x in st]) for st in states])
    out = engine.dbg_hades(buf, dense=dense)
    got = [[V.unmont(row[8 * i:8 * i + 8]) for i in range(5)] for row in out]
    assert got == [o.hades_perm(st) for st in states]


def test_fixed_base(engine):
    rnd = random.Random(15)
    ks = [0, 1, 2, 127, 128, 129, 255, 256, 0x8080, R - 1, R, R + 1, (1 << 252) - 1] + [rnd.randrange(R) for _ in range(100)]
    for base, B in ((0, o.G), (1, o.G_NUMS)):
        got = V.points_out(engine.dbg_scalar_mul(base, V.scalars(ks)))
        assert got == [V.mul(B, k) for k in ks]


@pytest.mark.parametrize("affine", [True, False])
def test_variable_base_whole_curve(engine, affine):
    """integer multiples on the WHOLE curve: identity, 8-torsion, points with a torsion component"""
    rnd = random.Random(16)
    pts = V.torsion_points() + [o.G, o.G_NUMS] + [V.rand_curve_point(rnd) for _ in range(12)]
    ks = [0, 1, 7, 8, 9, R - 1, R, R + 5, (1 << 250) - 1, (1 << 252) - 1] + [rnd.randrange(1 << 252) for _ in range(6)]
    P = [p for p in pts for _ in ks]
    K = [k for _ in pts for k in ks]
    zs = None if affine else [rnd.randrange(1, Q) for _ in P]
    got = V.points_out(engine.dbg_scalar_mul(2, V.scalars(K), V.points(P, zs), affine=affine))
    assert got == [V.mul(p, k) for p, k in zip(P, K)]


def make_single(rnd, n):
    sk = [rnd.randrange(R) for _ in range(n)]
    nonce = [rnd.randrange(R) for _ in range(n)]
    m = [rnd.randrange(Q) for _ in range(n)]
    # edge block: u = 0 (nonce = c*sk is unreachable; force via sk = 0, nonce = 0), m = 0, m = q-1, sk = 1, nonce = 1
    sk[:5] = [0, 1, R - 1, sk[3], sk[4]]
    nonce[:5] = [0, 1, R - 1, 1, nonce[4]]
    m[:5] = [0, Q - 1, 1, 0, Q - 1]
    return sk, nonce, m


def test_sign_and_keygen(engine):
    rnd = random.Random(17)
    n = 200
    sk, nonce, m = make_single(rnd, n)
    u, Rr, c = engine.sign(V.scalars(sk), V.fqs(m), V.scalars(nonce))
    exp = [o.sign(a, b, mm, mul=V.mul) for a, b, mm in zip(sk, nonce, m)]
    assert V.ints_out(u) == [e[0] for e in exp]
    assert V.points_out(Rr) == [e[1] for e in exp]
    assert V.ints_out(c) == [e[2] for e in exp]
    assert V.points_out(engine.keygen(V.scalars(sk))) == [V.mul(o.G, a) for a in sk]
    pk, pkp = engine.keygen_double(V.scalars(sk))
    assert V.points_out(pk) == [V.mul(o.G, a) for a in sk]
    assert V.points_out(pkp) == [V.mul(o.G_NUMS, a) for a in sk]


def test_sign_double_vargen(engine):
    rnd = random.Random(18)
    n = 96
    sk, nonce, m = make_single(rnd, n)
    u, Rr, Rp, c = engine.sign_double(V.scalars(sk), V.fqs(m), V.scalars(nonce))
    exp = [o.sign_double(a, b, mm, mul=V.mul) for a, b, mm in zip(sk, nonce, m)]
    assert V.ints_out(u) == [e[0] for e in exp]
    assert V.points_out(Rr) == [e[1] for e in exp]
    assert V.points_out(Rp) == [e[2] for e in exp]
    assert V.ints_out(c) == [e[3] for e in exp]
    gens = [V.mul(o.G, rnd.randrange(R)) for _ in range(n)]
    gens[0] = o.IDENTITY
    for affine in (True, False):
        zs = None if affine else [rnd.randrange(1, Q) for _ in gens]
        u, Rr, c = engine.sign_vargen(V.scalars(sk), V.points(gens, zs), V.fqs(m), V.scalars(nonce), affine=affine)
        exp = [o.sign_vargen(a, g, b, mm, mul=V.mul) for a, g, b, mm in zip(sk, gens, nonce, m)]
        assert V.ints_out(u) == [e[0] for e in exp]
        assert V.points_out(Rr) == [e[1] for e in exp]
        assert V.ints_out(c) == [e[2] for e in exp]
        assert V.points_out(engine.keygen_vargen(V.scalars(sk), V.points(gens, zs), affine=affine)) == \
            [V.mul(g, a) for g, a in zip(gens, sk)]


def corrupt_single(rnd, i, pk, u, Rp, m, pks):
    mode = i % 5
    if mode == 0:
        u = (u + 1) % R
    elif mode == 1:
        m = (m + 1) % Q
    elif mode == 2:
        Rp = o.pt_add(Rp, o.G)
    elif mode == 3:
        pk = pks[(i + 1) % len(pks)]
    else:
        pk = o.pt_add(pk, (0, Q - 1))  # add the order-2 point: same r-torsion part, wrong torsion component
    return pk, u, Rp, m


@pytest.mark.parametrize("affine", [True, False])
def test_verify_single(engine, affine):
    rnd = random.Random(19)
    n = 257  # ragged: not a multiple of 32
    sk, nonce, m = make_single(rnd, n)
    pks = [V.mul(o.G, a) for a in sk]
    sigs = [o.sign(a, b, mm, mul=V.mul) for a, b, mm in zip(sk, nonce, m)]
    tup = [(pks[i], sigs[i][0], sigs[i][1], m[i]) for i in range(n)]
    for i in range(n):
        if i % 3 == 2:
            tup[i] = corrupt_single(rnd, i // 3, *tup[i], pks)
    exp = [o.verify(*t, mul=V.mul) for t in tup]
    assert exp[:8] == [True, True, False, True, True, False, True, True]
    zs1 = None if affine else [rnd.randrange(1, Q) for _ in tup]
    zs2 = None if affine else [rnd.randrange(1, Q) for _ in tup]
    ok, c = engine.verify(V.points([t[0] for t in tup], zs1), V.scalars([t[1] for t in tup]),
                          V.points([t[2] for t in tup], zs2), V.fqs([t[3] for t in tup]), affine=affine)
    assert list(ok) == exp
    assert V.ints_out(c) == [o.challenge_hash(t[2], t[3]) for t in tup]


def test_verify_noncanonical_u_rejected(engine):
    rnd = random.Random(20)
    sk, nonce, m = rnd.randrange(R), rnd.randrange(R), rnd.randrange(Q)
    u, Rp, _ = o.sign(sk, nonce, m, mul=V.mul)
    pk = V.mul(o.G, sk)
    ok, _ = engine.verify(V.points([pk, pk]), V.scalars([u, u + R]), V.points([Rp, Rp]), V.fqs([m, m]))
    assert list(ok) == [True, False]  # JubJubScalar::from_bytes rejects u >= r


@pytest.mark.parametrize("affine", [True, False])
def test_verify_double(engine, affine):
    rnd = random.Random(21)
    n = 70
    sk, nonce, m = make_single(rnd, n)
    pk = [V.mul(o.G, a) for a in sk]
    pkp = [V.mul(o.G_NUMS, a) for a in sk]
    sigs = [o.sign_double(a, b, mm, mul=V.mul) for a, b, mm in zip(sk, nonce, m)]
    tup = [[pk[i], pkp[i], sigs[i][0], sigs[i][1], sigs[i][2], m[i]] for i in range(n)]
    for i in range(n):
        k = i % 7
        if k == 1: tup[i][2] = (tup[i][2] + 1) % R
        if k == 2: tup[i][1] = pkp[(i + 1) % n]           # only the second equation fails
        if k == 3: tup[i][3] = o.pt_add(tup[i][3], o.G)
        if k == 4: tup[i][4] = o.pt_add(tup[i][4], o.G_NUMS)
        if k == 5: tup[i][5] = (tup[i][5] + 1) % Q
    exp = [o.verify_double(*t, mul=V.mul) for t in tup]
    z = (lambda: None) if affine else (lambda: [rnd.randrange(1, Q) for _ in tup])
    ok, c = engine.verify_double(V.points([t[0] for t in tup], z()), V.points([t[1] for t in tup], z()),
                                 V.scalars([t[2] for t in tup]), V.points([t[3] for t in tup], z()),
                                 V.points([t[4] for t in tup], z()), V.fqs([t[5] for t in tup]), affine=affine)
    assert list(ok) == exp
    assert sum(exp) == sum(1 for i in range(n) if i % 7 in (0, 6))
    assert V.ints_out(c) == [o.challenge_hash_double(t[3], t[4], t[5]) for t in tup]


@pytest.mark.parametrize("affine", [True, False])
def test_verify_vargen(engine, affine):
    rnd = random.Random(22)
    n = 70
    sk, nonce, m = make_single(rnd, n)
    gens = [V.mul(o.G, rnd.randrange(R)) for _ in range(n)]
    pk = [V.mul(g, a) for g, a in zip(gens, sk)]
    sigs = [o.sign_vargen(a, g, b, mm, mul=V.mul) for a, g, b, mm in zip(sk, gens, nonce, m)]
    tup = [[pk[i], gens[i], sigs[i][0], sigs[i][1], m[i]] for i in range(n)]
    for i in range(n):
        k = i % 6
        if k == 1: tup[i][2] = (tup[i][2] + 1) % R
        if k == 2: tup[i][1] = gens[(i + 1) % n]
        if k == 3: tup[i][3] = o.pt_add(tup[i][3], o.G)
        if k == 4: tup[i][4] = (tup[i][4] + 1) % Q
        if k == 5: tup[i][0] = pk[(i + 1) % n]
    exp = [o.verify_vargen(*t, mul=V.mul) for t in tup]
    z = (lambda: None) if affine else (lambda: [rnd.randrange(1, Q) for _ in tup])
    ok, c = engine.verify_vargen(V.points([t[0] for t in tup], z()), V.points([t[1] for t in tup], z()),
                                 V.scalars([t[2] for t in tup]), V.points([t[3] for t in tup], z()),
                                 V.fqs([t[4] for t in tup]), affine=affine)
    assert list(ok) == exp
    assert V.ints_out(c) == [o.challenge_hash(t[3], t[4]) for t in tup]


def test_verify_small_order_keys(engine):
    """keys / R built with from_raw_unchecked may lie outside the prime-order subgroup; the verdict
    must still be the reference's integer-multiple verdict."""
    rnd = random.Random(23)
    tors = V.torsion_points()
    tup = []
    for t in tors:
        sk, nonce, m = rnd.randrange(R), rnd.randrange(R), rnd.randrange(Q)
        u, Rp, c = o.sign(sk, nonce, m, mul=V.mul)
        pk = o.pt_add(V.mul(o.G, sk), t)          # pk with a torsion component
        tup.append((pk, u, Rp, m))
        # R adjusted so the equation holds again: R' = uG + c'pk needs c' = H(R', m): not solvable in
        # general, so instead use pk = t alone with u = nonce (sk = 0): uG + c*t == R iff c*t == O
        tup.append((t, nonce, V.mul(o.G, nonce), m))
    exp = [o.verify(*t, mul=V.mul) for t in tup]
    ok, _ = engine.verify(V.points([t[0] for t in tup]), V.scalars([t[1] for t in tup]), V.points([t[2] for t in tup]),
                          V.fqs([t[3] for t in tup]))
    assert list(ok) == exp
    assert any(exp) and not all(exp)


def test_empty_batch(engine):
    ok, c = engine.verify(np.zeros((0, 16), np.uint32), np.zeros((0, 8), np.uint32), np.zeros((0, 16), np.uint32),
                          np.zeros((0, 8), np.uint32))
    assert ok.size == 0


# ---- wire formats on the device ---------------------------------------------------
def test_points_compress_decompress(engine):
    rnd = random.Random(30)
    pts = V.torsion_points() + [o.G, o.G_NUMS] + [V.rand_curve_point(rnd) for _ in range(120)]
    enc = b"".join(o.affine_to_bytes(p) for p in pts)
    got, ok = engine.points_decompress(enc)
    assert ok.all() and V.points_out(got) == pts
    assert engine.points_compress(V.points(pts)).tobytes() == enc
    zs = [rnd.randrange(1, Q) for _ in pts]
    assert engine.points_compress(V.points(pts, zs), affine=False).tobytes() == enc
    # what JubJubAffine::from_bytes rejects: non-canonical v, v not on the curve
    raw = [v.to_bytes(32, "little") for v in list(range(2, 200)) + [Q, Q + 1, (1 << 255) - 1]]
    exp = [o.affine_from_bytes(b) for b in raw]
    got, ok = engine.points_decompress(b"".join(raw))
    assert list(ok) == [e is not None for e in exp]
    assert [p for p, k in zip(V.points_out(got), ok) if k] == [e for e in exp if e is not None]
    assert not all(ok) and any(ok)
    # sign bit selects the root
    flipped = bytearray(o.affine_to_bytes(o.G)); flipped[31] ^= 0x80
    got, ok = engine.points_decompress(bytes(flipped))
    assert ok[0] and V.points_out(got)[0] == o.pt_neg(o.G)


def test_from_bytes_wide(engine):
    rnd = random.Random(31)
    ws = [0, 1, R, Q, (1 << 512) - 1, (1 << 256) - 1, 1 << 256, R << 256, Q << 200] + [rnd.randrange(1 << 512) for _ in range(300)]
    data = b"".join(w.to_bytes(64, "little") for w in ws)
    assert V.ints_out(engine.scalars_from_wide(data, 0)) == [w % R for w in ws]
    assert [V.unmont(r) for r in engine.scalars_from_wide(data, 1)] == [w % Q for w in ws]
    # = the draws of the reference's RNG: nonce i of a StdRng used only for signing
    rng = o.StdRng.seed_from_u64(2321)
    blocks = b"".join(o.chacha_block(rng.key, i) for i in range(8))
    assert V.ints_out(engine.scalars_from_wide(blocks, 0)) == [o.nonce_from_block(rng.key, i) for i in range(8)]
    xs = [rnd.randrange(Q) for _ in range(50)]
    assert V.ints_out(engine.fq_to_mont(V.scalars(xs))) == [x * V.RADIX % Q for x in xs]
    assert V.ints_out(engine.fq_from_mont(V.fqs(xs))) == xs


def test_sign_and_verify_bytes(engine):
    rnd = random.Random(32)
    n = 150
    sk, nonce, m = make_single(rnd, n)
    sigs = engine.sign_bytes(b"".join(x.to_bytes(32, "little") for x in sk), b"".join(x.to_bytes(32, "little") for x in m),
                             b"".join(x.to_bytes(32, "little") for x in nonce))
    exp = [o.sign(a, b, mm, mul=V.mul) for a, b, mm in zip(sk, nonce, m)]
    assert [bytes(s) for s in sigs] == [e[0].to_bytes(32, "little") + o.affine_to_bytes(e[1]) for e in exp]  # Signature::to_bytes
    pk = [o.affine_to_bytes(V.mul(o.G, a)) for a in sk]
    sig = [bytearray(s) for s in sigs]
    msg = [bytearray(x.to_bytes(32, "little")) for x in m]
    want_ok, want_inv = [True] * n, [False] * n
    for i in range(n):
        k = i % 8
        if k == 1:   # corrupted u (still canonical)
            sig[i][0] ^= 1; want_ok[i] = False
        elif k == 2:  # u >= r: JubJubScalar::from_bytes fails
            sig[i][:32] = (exp[i][0] + R).to_bytes(32, "little"); want_ok[i] = False; want_inv[i] = True
        elif k == 3:  # R bytes not on the curve
            bad = next(v for v in range(2 + i, 400) if o.affine_from_bytes(v.to_bytes(32, "little")) is None)
            sig[i][32:] = bad.to_bytes(32, "little"); want_ok[i] = False; want_inv[i] = True
        elif k == 4:  # message >= q
            msg[i][:] = (Q + i).to_bytes(32, "little"); want_ok[i] = False; want_inv[i] = True
        elif k == 5:  # someone else's key
            pk[i] = pk[(i + 1) % n]; want_ok[i] = False
        elif k == 6:  # pk bytes with non-canonical v
            pk[i] = (Q + 3).to_bytes(32, "little"); want_ok[i] = False; want_inv[i] = True
    ok, inv = engine.verify_bytes(b"".join(pk), b"".join(bytes(s) for s in sig), b"".join(bytes(x) for x in msg))
    assert list(ok) == want_ok and list(inv) == want_inv
