"""Opcode histogram (executed warp-instructions, stall samples) from `ncu --page source --csv`.
usage: ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_opmix.py src.csv"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ex, smp = collections.Counter(), collections.Counter()
n_static = 0
for r in rows[2:]:
    if len(r) <= iex: continue
    toks = r[isrc].split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.rstrip(";")
    try:
        ex[op] += int(r[iex]); smp[op] += int(r[ismp]); n_static += 1
    except ValueError:
        pass
tot, tots = sum(ex.values()), sum(smp.values())
print(f"static instructions {n_static}  ({n_static*16/1024:.0f} KB)   executed warp-instr {tot}   samples {tots}")
for op, c in ex.most_common(22):
    print(f"{op:22s} {c:14d} {100*c/tot:6.2f}%   samples {100*smp[op]/max(tots,1):6.2f}%")
