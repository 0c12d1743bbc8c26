"""profiles/op_counts.json (the roofline's "work per tuple") must equal what the instrumented host build of the
per-tuple code counts today -- so a kernel change cannot silently leave a stale roofline denominator."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_op_counts_file_is_current():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "op_counts.py"), "--check"], capture_output=True,
                         text=True, timeout=900)
    assert out.returncode == 0, out.stderr
    now = json.loads(out.stdout)
    saved = json.load(open(os.path.join(ROOT, "profiles", "op_counts.json")))
    for op, v in now.items():
        if isinstance(v, dict):
            for k in ("imad_wide_per_tuple", "fq_mul", "fq_sqr", "fq_dot5"):
                assert saved[op][k] == v[k], (op, k)
    v = now["verify_affine"]
    field = 120 * v["fq_mul"] + 84 * v["fq_sqr"] + 368 * v["fq_dot5"] + 128 * v["fr_mont_mul"]
    # the rest is the half-size-scalar Euclid (13 limb products per step, ~75-110 steps)
    assert 0 <= v["imad_wide_per_tuple"] - field <= 13 * 160
