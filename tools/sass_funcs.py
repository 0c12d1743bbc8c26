#!/usr/bin/env python3
"""Per-device-function opcode tally of one kernel in the built library (development aid).
usage: tools/sass_funcs.py <kernel-substring> [lib]     e.g.  tools/sass_funcs.py k_runILi11E"""
import collections, re, subprocess, sys
key = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "schnorr_b200/libschnorr_b200.so"
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
subs = []
for l in elf.split("\n"):
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+\S+\s+\S+\s+\S+\s+\$(\S*%s\S*?)\$(\S+)" % re.escape(key), l)
    if m:
        subs.append((int(m.group(1), 16), int(m.group(2), 16), m.group(4)))
names = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = names.split("\t\tFunction : ")
blk = next(b for b in blocks if key in b.split("\n")[0])
tally = collections.defaultdict(collections.Counter)
for l in blk.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
    if not m:
        continue
    addr = int(m.group(1), 16)
    owner = "<kernel body>"
    for off, size, nm in subs:
        if off <= addr < off + size:
            owner = re.sub(r"^_ZN\d+_INTERNAL_[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+\d", "", nm)[:48]
    tally[owner][m.group(2)] += 1
for nm, c in sorted(tally.items(), key=lambda kv: -sum(kv[1].values())):
    print("%-50s %5d instr" % (nm, sum(c.values())))
    print("     " + ", ".join("%s:%d" % kv for kv in c.most_common(14)))
