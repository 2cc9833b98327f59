#!/usr/bin/env python
"""Quick look at one .ncu-rep: headline metrics, stall reasons, opcode histogram.  usage: python tools/ncu_quick.py rep [launch]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
hdr, units = raw[0], raw[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts.sum"]
for vals in raw[2:]:
    print("==", vals[hdr.index("Kernel Name")])
    for i, h in enumerate(hdr):
        if h in keys:
            print(f"  {h} [{units[i]}] = {vals[i]}")
    st = [(float(vals[i]), h.split("issue_stalled_")[1].split("_per_")[0]) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    print("  stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:7]))
src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
h2 = src[1]
iS, iI, iW = h2.index("Source"), h2.index("Instructions Executed"), h2.index("# Samples")
hist, samp, tot = collections.Counter(), collections.Counter(), 0
for r in src[2:]:
    if len(r) < len(h2):
        continue
    try:
        n, s = int(r[iI]), int(r[iW])
    except ValueError:
        continue
    t = r[iS].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    hist[op] += n
    samp[op] += s
    tot += n
print("  opcodes:", ", ".join(f"{k}={100 * v / tot:.1f}%({samp[k]})" for k, v in hist.most_common(16)))
