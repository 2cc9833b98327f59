#!/bin/bash
# round 2, second GPU call (1 GPU): memcheck of the failing large test, whole GPU suite, AMG sweeps with the fused coarse kernel, launch list of a cfg4 Newton step
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cp gpurun_out/newton_samples_cfg4_193.json tests/golden/ 2>/dev/null
( timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_large.py -x -q -k "cfg3_counts or cfg4" > gpurun_out/r2_run2_memcheck.log 2>&1 ; tail -30 gpurun_out/r2_run2_memcheck.log )
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_run2_pytest.log 2>&1
tail -15 gpurun_out/r2_run2_pytest.log
for W in 0 1 2 3 6; do
WL=cfg3 METHODS="cg+amg+WDEPTH=$W" timeout 300 python tools/linsolve_probe.py 2>&1 | tail -1
done | tee gpurun_out/r2_run2_sweep_cfg3.log
VFVM_AMG_NO_FUSE=1 WL=cfg3 METHODS="cg+amg+WDEPTH=0,cg+amg+WDEPTH=2" timeout 300 python tools/linsolve_probe.py 2>&1 | tail -2 | tee -a gpurun_out/r2_run2_sweep_cfg3.log
WL=cfg3 METHODS="cg+amg+WDEPTH=6+COARSE_SWEEPS=8,cg+amg+WDEPTH=6+SWEEPS=2,cg+amg+WDEPTH=6+ALPHA=1.5,cg+amg+WDEPTH=6+ALPHA=2.0,cg+amg+WDEPTH=6+OMEGA=0.7,cg+amg+WDEPTH=6+OMEGA=0.9" timeout 600 python tools/linsolve_probe.py 2>&1 | tail -6 | tee -a gpurun_out/r2_run2_sweep_cfg3.log
for W in 0 2 6; do
WL=cfg4 METHODS="bicgstab+amg+WDEPTH=$W" timeout 300 python tools/linsolve_probe.py 2>&1 | tail -1
done | tee gpurun_out/r2_run2_sweep_cfg4.log
VFVM_AMG_NO_FUSE=1 WL=cfg4 METHODS="bicgstab+amg+WDEPTH=0" timeout 300 python tools/linsolve_probe.py 2>&1 | tail -1 | tee -a gpurun_out/r2_run2_sweep_cfg4.log
WL=cfg4 METHODS="gmres+amg+WDEPTH=2" timeout 300 python tools/linsolve_probe.py 2>&1 | tail -1 | tee -a gpurun_out/r2_run2_sweep_cfg4.log
# launch list (time per launch) of one cfg4 Newton step: which kernels dominate the 15 ms per BiCGStab iteration
WL=cfg4 NX=129 METHODS="bicgstab+amg+WDEPTH=0" MAXIT=6 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_run2_launches_cfg4_129.csv python tools/linsolve_probe.py > gpurun_out/r2_run2_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r2_run2_launches_cfg4_129.csv 2>&1 | tail -40 | tee gpurun_out/r2_run2_launches_cfg4_129_summary.txt
