#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv):
launch count, total / average time and DRAM GB/s per kernel, split into launches that move more / less than 20 MB.
usage: python tools/launch_summary.py launches.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
iK, iM, iV, iU, iID = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
L = collections.OrderedDict()
for r in rows[1:]:
    name = r[iK].replace("<unnamed>::", "").replace("(anonymous namespace)::", "").replace("void ", "").replace("(int)", "").replace("(bool)", "")
    d = L.setdefault(r[iID], {"k": name.split("(")[0]})
    v = float(r[iV].replace(",", ""))
    if "duration" in r[iM]:
        d["ns"] = v * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r[iU], 1)
    else:
        d["bytes"] = d.get("bytes", 0) + v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[iU], 1)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for d in L.values():
    key = (d["k"], ">20MB" if d.get("bytes", 0) > 2e7 else "small")
    agg[key][0] += 1
    agg[key][1] += d["ns"]
    agg[key][2] += d.get("bytes", 0)
tot = sum(v[1] for v in agg.values())
print(f"{len(L)} launches, {tot / 1e3:.1f} us of kernel time (serialised under ncu)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[0]:50s} {k[1]:6s} n={v[0]:4d}  total {v[1] / 1e3:8.1f} us ({100 * v[1] / tot:4.1f} %)  avg {v[1] / v[0] / 1e3:7.1f} us  DRAM {v[2] / max(v[1], 1):7.1f} GB/s")
