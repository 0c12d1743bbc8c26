"""GPU vs host build of csrc/lat3.cuh on the same (c, u): prints where they differ (development aid)."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import schnorr_oracle as o
import vectors as V
import hostlib as H
from schnorr_b200 import Engine
R, N = o.R, 8 * o.R
rnd = random.Random(5)
cases = [(rnd.randrange(1 << 250), rnd.randrange(R)) for _ in range(64)]
e = Engine([0])
out = e.dbg_lattice3(V.scalars([c for c, _ in cases]), V.scalars([u for _, u in cases]))
lib = H.build()
bad = 0
for k, ((c, u), row) in enumerate(zip(cases, out)):
    a, b, d = V.to_int(row[:8]), V.to_int(row[8:16]), V.to_int(row[16:24])
    a, b, d = (-a if row[24] else a), (-b if row[25] else b), (-d if row[26] else d)
    h = np.zeros(28, np.uint32)
    lib.h_lattice3(H.ptr(H.limbs(c)), H.ptr(H.limbs(u)), H.ptr(h))
    ha, hb, hd = H.to_int(h[:8]), H.to_int(h[8:16]), H.to_int(h[16:24])
    ha, hb, hd = (-ha if h[24] else ha), (-hb if h[25] else hb), (-hd if h[26] else hd)
    rel = (a - b * c) % N == 0 and (d - b * u) % N == 0
    same = (a, b, d, int(row[27])) == (ha, hb, hd, int(h[27]))
    if not rel or not same:
        bad += 1
        if bad <= 6:
            print(f"case {k}: ok={row[27]} rel={rel} same_as_host={same} bits(a,b,d)={abs(a).bit_length()},{abs(b).bit_length()},{abs(d).bit_length()} "
                  f"host ok={h[27]} bits={abs(ha).bit_length()},{abs(hb).bit_length()},{abs(hd).bit_length()}")
            print("   gpu  b=%x a=%x" % (b, a)); print("   host b=%x a=%x" % (hb, ha))
print("cases", len(cases), "bad", bad)
