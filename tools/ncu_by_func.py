#!/usr/bin/env python3
"""Per-device-function totals (executed warp-instructions, samples, stall reasons) from an ncu source page.
usage: ncu -i X.ncu-rep --page source --csv > src.csv; tools/ncu_by_func.py src.csv <lib.so> <kernel-substring>"""
import collections, csv, re, subprocess, sys
src, lib, key = sys.argv[1:4]
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
subs = []
for l in elf.split("\n"):
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+\S+\s+\S+\s+\S+\s+\$(\S*%s\S*?)\$(\S+)" % re.escape(key), l)
    if m:
        subs.append((int(m.group(1), 16), int(m.group(2), 16), re.sub(r"^_ZN\d+_INTERNAL_[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+\d", "", m.group(4))[:40]))
rows = list(csv.reader(open(src, errors="replace")))
hdr = rows[1]
ia, iex, ismp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
agg = collections.defaultdict(lambda: collections.Counter())
for r in rows[2:]:
    if len(r) <= ismp or not r[ia] or r[ia] == "Address":
        continue
    addr = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    if base is None:
        base = addr
    off = addr - base
    owner = "<kernel body>"
    for o, sz, nm in subs:
        if o <= off < o + sz:
            owner = nm
    a = agg[owner]
    try:
        a["instr"] += int(r[iex]); a["samples"] += int(r[ismp])
    except ValueError:
        continue
    for i, h in stall_cols:
        try:
            a[h] += int(r[i])
        except ValueError:
            pass
tot = sum(a["samples"] for a in agg.values())
for nm, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
    st = sorted(((v, k) for k, v in a.items() if k.startswith("stall_")), reverse=True)[:6]
    print("%-42s instr %12d  samples %8d (%5.1f%%)  %s" % (nm, a["instr"], a["samples"], 100.0 * a["samples"] / max(tot, 1),
          " ".join("%s=%.0f%%" % (k[6:], 100.0 * v / max(a["samples"], 1)) for v, k in st)))
