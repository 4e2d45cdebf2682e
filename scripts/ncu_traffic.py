#!/usr/bin/env python
"""DRAM traffic per launch of the dominant kernels from an `ncu --set full` raw CSV -> profiles/<round>/traffic.json
usage: ncu_traffic.py raw.csv out.json  (bench.py reads roofline.traffic from the json)"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
kn = hdr.index("Kernel Name")
def col(name):
    return hdr.index(name)
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for r in rows[2:]:
    name = r[kn].split("(")[0].replace("void ", "").strip()
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = col(m)
        tot += float(r[i].replace(",", "")) * scale[units[i]]
    out.setdefault(name, []).append(tot)
res = {k: {"dram_bytes_per_launch": sum(v) / len(v), "launches": len(v)} for k, v in out.items()}
res["_source"] = "ncu --set full --clock-control none (dram__bytes_read.sum + dram__bytes_write.sum), C2 workload, " + sys.argv[1]
json.dump(res, open(sys.argv[2], "w"), indent=1, sort_keys=True)
print(json.dumps(res, indent=1, sort_keys=True))
