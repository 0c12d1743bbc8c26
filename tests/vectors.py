"""Seeded synthetic tuples for the parity tests, produced with the oracle (CPU)."""
import random

import numpy as np

import schnorr_oracle as o

Q, R = o.Q, o.R
RADIX = 1 << 256
RINV = pow(RADIX, -1, Q)


def limbs(x):
    return np.frombuffer(int(x).to_bytes(32, "little"), dtype=np.uint32)


def to_int(a):
    return int.from_bytes(np.ascontiguousarray(a, dtype=np.uint32).tobytes(), "little")


def mont(x):
    return limbs(x * RADIX % Q)


def unmont(a):
    return to_int(a) * RINV % Q


def scalars(xs):
    return np.stack([limbs(x) for x in xs]).copy()


def fqs(xs):
    return np.stack([mont(x) for x in xs]).copy()


def points(ps, zs=None):
    """affine points -> [n,16] Montgomery limbs, or [n,24] projective (U,V,Z)=(u z, v z, z) when zs is given"""
    rows = []
    for i, p in enumerate(ps):
        if zs is None:
            rows.append(np.concatenate([mont(p[0]), mont(p[1])]))
        else:
            z = zs[i]
            rows.append(np.concatenate([mont(p[0] * z % Q), mont(p[1] * z % Q), mont(z)]))
    return np.stack(rows).copy()


def points_out(a):
    a = np.asarray(a).reshape(-1, 16)
    return [(unmont(r[:8]), unmont(r[8:])) for r in a]


def ints_out(a):
    return [to_int(r) for r in np.asarray(a).reshape(-1, 8)]


def rand_curve_point(rnd):
    """random point on the whole curve (cofactor NOT cleared: may carry an 8-torsion component)"""
    while True:
        p = o.affine_from_bytes(rnd.randrange(Q).to_bytes(32, "little"))
        if p is not None:
            return p


def torsion_points():
    """the points of order 1, 2, 4 and 8 reachable as r * P"""
    rnd = random.Random(99)
    out = {o.IDENTITY, (0, Q - 1)}
    while len(out) < 8:
        t = o.pt_mul_fast(rand_curve_point(rnd), R)
        k = t
        for _ in range(8):
            out.add(k)
            k = o.pt_add(k, t)
    return sorted(out)


mul = o.pt_mul_fast


def hgcd_hostile_challenges(count, seed=5):
    """Challenges c < 2^250 for which the half-size-scalar path (csrc/hgcd.cuh) must give up and the kernel falls back
    to the full-size multiplication: Euclid on (8r, c) has consecutive remainders (A, s) with A >= 2^134, s < 2^121 and
    an EVEN cofactor T on s, so neither (s, T) nor its neighbours fit the 134-bit window budget with an odd b.
    Built backwards from A T + s T' = 8r  (T' odd, coprime to T):  c = -A / T' mod 8r."""
    import math
    N = 8 * R
    rnd, out = random.Random(seed), []
    while len(out) < count:
        T = rnd.randrange(1 << 116, 1 << 119) & ~1
        Tp = rnd.randrange(1, T) | 1
        if math.gcd(T, Tp) != 1 or math.gcd(Tp, N) != 1:
            continue
        s = N * pow(Tp, -1, T) % T
        A = (N - s * Tp) // T
        c = (-A * pow(Tp, -1, N)) % N
        if c < (1 << 250):
            out.append(c)
    return out
