#!/usr/bin/env python3
"""SASS evidence for profiles/ (SURVEY.md 8(d)): from the built library, for one kernel
  * the per-device-function opcode tally (what the hot code is made of),
  * the full body of the out-of-line field multiplier (the IMAD.WIDE.U32[.X] carry chains), and
  * the innermost loop of the kernel body that holds the doubling chain (its CALL sequence and the argument moves
    ptxas places around the calls).
usage: tools/sass_excerpt.py <kernel-substring> [lib] > profiles/sass_<name>.txt      e.g. k_runILi19E"""
import collections, re, subprocess, sys

key = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "schnorr_b200/libschnorr_b200.so"
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
subs = []
for l in elf.split("\n"):
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+\S+\s+\S+\s+\S+\s+\$(\S*%s\S*?)\$(\S+)" % re.escape(key), l)
    if m:
        subs.append((int(m.group(1), 16), int(m.group(2), 16), re.sub(r"^_ZN\d+_INTERNAL_[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+\d", "", m.group(4))))
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blk = next(b for b in sass.split("\t\tFunction : ") if key in b.split("\n")[0])
ins = []  # (addr, text)
for l in blk.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?)\s*;", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))


def owner(addr):
    for off, size, nm in subs:
        if off <= addr < off + size:
            return nm
    return "<kernel body>"


def opcode(t):
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    return t.split()[0]


tally = collections.defaultdict(collections.Counter)
for a, t in ins:
    tally[owner(a)][opcode(t)] += 1
print(f"# {key} in {lib}: {len(ins)} instructions ({len(ins) * 16 // 1024} KB)")
print("# ---- per device function ----")
for nm, c in sorted(tally.items(), key=lambda kv: -sum(kv[1].values())):
    print("# %-44s %5d instr   %s" % (nm[:44], sum(c.values()), ", ".join("%s:%d" % kv for kv in c.most_common(8))))
mul = next((s for s in subs if "fq_mul_ool" in s[2]), None)
if mul:
    print(f"\n# ---- {mul[2]}: the out-of-line Montgomery multiplier, {mul[1] // 16} instructions ----")
    for a, t in ins:
        if mul[0] <= a < mul[0] + mul[1]:
            print("  /*%05x*/ %s" % (a, t))
# innermost loop of the kernel body with the most CALLs: backward branches = loops
body = [(a, t) for a, t in ins if owner(a) == "<kernel body>"]
loops = []
for a, t in body:
    m = re.search(r"\bBRA\S*\s+.*?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        lo = int(m.group(1), 16)
        calls = sum(1 for x, u in body if lo <= x <= a and "CALL" in u)
        loops.append((a - lo, lo, a, calls))
cands = [l for l in loops if l[3] >= 7]
if cands:
    size, lo, hi, calls = min(cands)
    print(f"\n# ---- innermost kernel-body loop with >= 7 calls: {size // 16 + 1} instructions, {calls} CALLs (the rolled doubling) ----")
    for a, t in body:
        if lo <= a <= hi:
            print("  /*%05x*/ %s" % (a, t))
