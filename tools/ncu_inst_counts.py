#!/usr/bin/env python3
"""Executed warp instructions per tuple of an op, from the ncu summaries under profiles/ (tools/ncu_summary.py output):
sum over the op's kernels of smsp__inst_executed.sum / (grid x 128 threads / 32).  Writes profiles/inst_counts.json, which
bench.py joins with profiles/op_counts.json (IMAD.WIDE per tuple) for the issue-model record.
usage: tools/ncu_inst_counts.py <op-key>=<summary.txt>[@tuples-per-thread | #tuples-in-the-launch] ...
(#tuples for persistent kernels, whose grid does not tell the tuple count)"""
import json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(ROOT, "profiles", "inst_counts.json")
out = json.load(open(path)) if os.path.exists(path) else {}
for arg in sys.argv[1:]:
    key, f = arg.split("=")
    f, _, ntup = f.partition("#")
    f, _, tpt = f.partition("@")
    tpt = float(tpt or 1)
    ntup = float(ntup) if ntup else None
    kern, total, cyc = [], 0.0, 0.0
    name = grid = None
    seen = set()
    for line in open(f):
        t = line.split()
        if line.startswith("Kernel Name"):
            name = line[len("Kernel Name"):].strip()
        elif t and t[0] == "launch__grid_size":
            grid = float(t[1])
        elif t and t[0] == "smsp__inst_executed.sum" and name not in seen:
            seen.add(name)  # the first launch of each kernel
            per_tuple = float(t[1]) * 32.0 / (ntup if ntup else grid * 128.0 * tpt)
            kern.append({"kernel": name, "warp_inst_per_tuple": round(per_tuple, 1)})
            total += per_tuple
    out[key] = {"warp_inst_per_tuple": round(total, 1), "kernels": kern, "source": os.path.basename(f)}
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
