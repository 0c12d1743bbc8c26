#!/usr/bin/env python3
"""Static instruction counts of the main loop of every kernel in build/issue_mix (from cuobjdump -sass): the loop is the span of the
longest backward branch.  Prints JSON {"KIND,K": {"wide": n, "other": n, "by_op": {...}, "iterations_unrolled": u}} -- joined with the
measured cycles by tools/issue_mix_fit.py."""
import collections, json, re, subprocess, sys
exe = sys.argv[1] if len(sys.argv) > 1 else "build/issue_mix"
sass = subprocess.run(["cuobjdump", "-sass", exe], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = funcs.setdefault(m.group(1), [])
        continue
    g = re.match(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if g and cur is not None:
        cur.append((int(g.group(1), 16), g.group(2).strip()))
out = {}
for f, ins in funcs.items():
    k = re.search(r"ILi(\d+)ELi(\d+)E", f)
    best = None
    for addr, txt in ins:
        m = re.search(r"\bBRA\b.* (0x[0-9a-f]+)$", txt)
        if m and int(m.group(1), 16) < addr:
            span = (int(m.group(1), 16), addr)
            if best is None or span[1] - span[0] > best[1] - best[0]:
                best = span
    if not best or not k:
        continue
    ops = collections.Counter()
    for addr, txt in ins:
        if best[0] <= addr <= best[1]:
            t = txt.split()
            mn = t[1] if t[0].startswith("@") else t[0]
            ops[mn] += 1
    wide = sum(v for o, v in ops.items() if o.startswith("IMAD.WIDE"))
    out[f"{k.group(1)},{k.group(2)}"] = {"wide": wide, "other": sum(ops.values()) - wide, "iterations_unrolled": wide // 8, "by_op": dict(ops.most_common())}
json.dump(out, sys.stdout, indent=1)
