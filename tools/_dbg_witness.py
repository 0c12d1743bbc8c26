import sys, random
sys.path[:0] = [".", "oracle", "tests"]
import numpy as np
import schnorr_oracle as o, vectors as V
from schnorr_b200 import Engine
e = Engine([0])
rnd = random.Random(60)
for n in (1, 37):
    sk = [rnd.randrange(1, o.R) for _ in range(n)]; nonce = [rnd.randrange(o.R) for _ in range(n)]; msg = [rnd.randrange(o.Q) for _ in range(n)]
    rows = e.sign_witness(0, V.scalars(sk), V.fqs(msg), V.scalars(nonce))
    for i in (0, n - 1):
        got = [V.unmont(rows[i, k]) for k in range(11)]
        u, Rp, c = o.sign(sk[i], nonce[i], msg[i], mul=V.mul)
        pk = V.mul(o.G, sk[i]); sa, sb = V.mul(o.G, u), V.mul(pk, c)
        want = [u, *Rp, *pk, msg[i], c, *sa, *sb]
        print(n, i, [g == w for g, w in zip(got, want)])
        print("   raw row0 limbs", rows[i, 0], "got0", hex(got[0]))
    u2, R2, c2 = e.sign(V.scalars(sk), V.fqs(msg), V.scalars(nonce))
    print("sign ok:", V.ints_out(u2)[0] == o.sign(sk[0], nonce[0], msg[0], mul=V.mul)[0])
