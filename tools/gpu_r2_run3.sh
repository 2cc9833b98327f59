#!/bin/bash
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( echo "--- default"; timeout 120 python tools/diag_bipolar.py
  echo "--- EVERY=1 (rebuild the preconditioner every Newton step)"; EVERY=1 timeout 120 python tools/diag_bipolar.py
  echo "--- NO_FUSE"; VFVM_AMG_NO_FUSE=1 timeout 120 python tools/diag_bipolar.py
  echo "--- NO_KRYLOV_GRAPH"; VFVM_NO_KRYLOV_GRAPH=1 timeout 120 python tools/diag_bipolar.py
  echo "--- NO_FUSE NO_GRAPH"; VFVM_AMG_NO_FUSE=1 VFVM_NO_KRYLOV_GRAPH=1 timeout 120 python tools/diag_bipolar.py
  echo "--- SPMV_BULK=0"; VFVM_SPMV_BULK=0 timeout 120 python tools/diag_bipolar.py ) > gpurun_out/r2_run3_diag.log 2>&1
grep -v "newton" gpurun_out/r2_run3_diag.log | tail -30
( timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_large.py -x -q -k "cfg3_counts or cfg4_full" > gpurun_out/r2_run3_memcheck.log 2>&1 ; tail -30 gpurun_out/r2_run3_memcheck.log )
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_run3_pytest.log 2>&1
tail -8 gpurun_out/r2_run3_pytest.log
for W in 0 2 6; do
WL=cfg3 METHODS="cg+amg+WDEPTH=$W" timeout 300 python tools/linsolve_probe.py 2>&1 | tail -1
done | tee gpurun_out/r2_run3_sweep_cfg3.log
VFVM_SPMV_BULK=2 WL=cfg3 METHODS="cg+amg+WDEPTH=2" timeout 300 python tools/linsolve_probe.py 2>&1 | tail -1 | tee -a gpurun_out/r2_run3_sweep_cfg3.log
WL=cfg4 METHODS="bicgstab+amg+WDEPTH=0,bicgstab+amg+WDEPTH=2" timeout 300 python tools/linsolve_probe.py 2>&1 | tail -2 | tee gpurun_out/r2_run3_sweep_cfg4.log
VFVM_SPMV_BULK=0 WL=cfg4 METHODS="bicgstab+amg+WDEPTH=0" timeout 300 python tools/linsolve_probe.py 2>&1 | tail -1 | tee -a gpurun_out/r2_run3_sweep_cfg4.log
WL=cfg5 METHODS="cg+amg+WDEPTH=0" timeout 300 python tools/linsolve_probe.py 2>&1 | tail -1 | tee gpurun_out/r2_run3_sweep_cfg5.log
WL=cfg2 METHODS="bicgstab+amg+WDEPTH=0,bicgstab+amg+WDEPTH=3" timeout 300 python tools/linsolve_probe.py 2>&1 | tail -2 | tee gpurun_out/r2_run3_sweep_cfg2.log
# full ncu capture of the two SpMV kernels for coupled species (one launch each)
WL=cfg4 NX=129 METHODS="bicgstab+block" MAXIT=4 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_spmv -s 2 -c 1 -o gpurun_out/r2_run3_spmv_bulk_cfg4 -f python tools/linsolve_probe.py > gpurun_out/r2_run3_ncu1.log 2>&1
VFVM_SPMV_BULK=0 WL=cfg4 NX=129 METHODS="bicgstab+block" MAXIT=4 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_spmv -s 2 -c 1 -o gpurun_out/r2_run3_spmv_regs_cfg4 -f python tools/linsolve_probe.py > gpurun_out/r2_run3_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
