#!/usr/bin/env python
"""Turns a .ncu-rep into the small per-kernel metric table that is committed under profiles/.
usage: python tools/ncu_summarize.py in.ncu-rep out.csv"""
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput", "gpu__dram_throughput", "sm__warps_active", "launch__",
        "sm__throughput", "hit_rate", "smsp__average_warps_issue_stalled", "smsp__inst_executed.sum", "smsp__issue_active", "smsp__warps_eligible", "pipe_fp64",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts.sum", "lts__throughput", "l1tex__throughput", "Kernel Name"]


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    keep = [h for h in hdr if any(s in h for s in KEEP)]
    with open(out, "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(rows) - 2)])
        for k in keep:
            i = hdr.index(k)
            w.writerow([k, units[i]] + [r[i] for r in rows[2:]])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
