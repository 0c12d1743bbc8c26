#!/usr/bin/env python3
"""Extract DRAM traffic of one kernel launch from an .ncu-rep into profiles/ncu_traffic.json.
usage: tools/ncu_traffic.py <op-name> <tuples-in-that-launch> X.ncu-rep [out.json]   (default profiles/ncu_traffic.json)"""
import csv, io, json, os, subprocess, sys
op, tuples, rep = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
recs = [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr)]  # one row per captured launch: summed (a step may be two launches)
d = {"Kernel Name": " + ".join(r["Kernel Name"] for r in recs),
     "gpu__time_duration.sum": str(sum(float(r["gpu__time_duration.sum"].replace(",", "")) for r in recs))}
def val(k):
    u = units[hdr.index(k)]
    return sum(float(r[k].replace(",", "")) for r in recs) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
path = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
out = json.load(open(path)) if os.path.exists(path) else {}
out[op] = {"kernel": d["Kernel Name"], "tuples": tuples, "dram_bytes_read": val("dram__bytes_read.sum"),
           "dram_bytes_write": val("dram__bytes_write.sum"), "gpu_time_ms": float(d["gpu__time_duration.sum"].replace(",", "")) *
           {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(units[hdr.index("gpu__time_duration.sum")], 1), "source": os.path.basename(rep)}
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out[op]))
