#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2_run4_pytest.log 2>&1
tail -12 gpurun_out/r2_run4_pytest.log
for W in 0 2; do
WL=cfg3 METHODS="cg+amg+WDEPTH=$W" timeout 300 python tools/linsolve_probe.py 2>&1 | tail -1
VFVM_AMG_NO_FUSE=1 WL=cfg3 METHODS="cg+amg+WDEPTH=$W" timeout 300 python tools/linsolve_probe.py 2>&1 | tail -1
done | tee gpurun_out/r2_run4_sweep_cfg3.log
WL=cfg4 METHODS="bicgstab+amg+WDEPTH=0" timeout 300 python tools/linsolve_probe.py 2>&1 | tail -1 | tee gpurun_out/r2_run4_sweep_cfg4.log
VFVM_AMG_NO_FUSE=1 WL=cfg4 METHODS="bicgstab+amg+WDEPTH=0" timeout 300 python tools/linsolve_probe.py 2>&1 | tail -1 | tee -a gpurun_out/r2_run4_sweep_cfg4.log
WL=cfg4 NX=129 METHODS="bicgstab+amg+WDEPTH=0" MAXIT=6 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_run4_launches_cfg4_129.csv python tools/linsolve_probe.py > gpurun_out/r2_run4_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r2_run4_launches_cfg4_129.csv 2>&1 | head -14 | tee gpurun_out/r2_run4_launches_cfg4_129_summary.txt
