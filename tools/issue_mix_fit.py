#!/usr/bin/env python3
"""Join the measured cycles of build/issue_mix (JSON on stdin or argv[1]) with the static loop contents (tools/issue_mix_static.py) and
fit   cycles per warp-iteration = a * (IMAD.WIDE count) + b * (other instructions)   per kind and over all points.
usage: build/issue_mix > m.json; python tools/issue_mix_fit.py m.json [build/issue_mix]"""
import json, subprocess, sys
import numpy as np
meas = json.load(open(sys.argv[1]))
exe = sys.argv[2] if len(sys.argv) > 2 else "build/issue_mix"
stat = json.loads(subprocess.run([sys.executable, __file__.replace("issue_mix_fit", "issue_mix_static"), exe], capture_output=True, text=True).stdout)
kinds = []
for p in meas["points"]:
    if p["kind"] not in kinds:
        kinds.append(p["kind"])
rows = []
for p in meas["points"]:
    s = stat[f"{kinds.index(p['kind'])},{p['extra_per_8_wide']}"]
    u = s["iterations_unrolled"]
    rows.append((p["kind"], s["wide"] / u, s["other"] / u, p["cycles_per_warp_iteration"], {k: v / u for k, v in s["by_op"].items() if not k.startswith("IMAD.WIDE")}))
out = {"gpu": meas["gpu"], "model": "cycles per warp-iteration on one SM sub-partition (4 warps resident) = a * wide + b * other", "kinds": {}}
def fit(rs):
    A = np.array([[r[1], r[2]] for r in rs]); y = np.array([r[3] for r in rs])
    (a, b), res, *_ = np.linalg.lstsq(A, y, rcond=None)
    pred = A @ np.array([a, b])
    return {"a_cycles_per_wide": round(float(a), 3), "b_cycles_per_other": round(float(b), 3), "max_rel_err": round(float(np.max(np.abs(pred - y) / y)), 3)}
for k in kinds:
    rs = [r for r in rows if r[0] == k]
    out["kinds"][k] = {"fit": fit(rs), "points": [{"wide": r[1], "other": r[2], "cycles": r[3], "other_ops": r[4]} for r in rs]}
out["all"] = fit(rows)
json.dump(out, sys.stdout, indent=1)
print()
