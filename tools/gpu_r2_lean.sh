#!/bin/bash
# 2 GPUs: masked-system diagnosis (replication on / off), lean coarse-level kernels on one GPU (A/B in one process) and on two.
# (Record of a measurement: the "lean" kernels it switched with LEAN= were slower and have been removed again, profiles/r2_amg_sweeps.txt.)
N=${1:-2}
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
VFVM_AMG_VERBOSE=1 timeout 240 $TR --master-port 29512 tests/mgpu_check.py > gpurun_out/r2_lean_mgpu_check_$N.log 2>&1
grep "mgpu_check\|MGPU_OK\|Error\|error" gpurun_out/r2_lean_mgpu_check_$N.log | cut -c1-900 | tail -12
MGPU_ONLY=masked VFVM_AMG_REPL_MAX_N=0 timeout 120 $TR --master-port 29516 tests/mgpu_check.py > gpurun_out/r2_lean_mgpu_check_${N}_norepl.log 2>&1
grep "mgpu_check\|MGPU_OK\|Error\|error" gpurun_out/r2_lean_mgpu_check_${N}_norepl.log | cut -c1-900 | tail -6
MGPU_ONLY=masked VFVM_AMG_LEAN=0 timeout 120 $TR --master-port 29517 tests/mgpu_check.py > gpurun_out/r2_lean_mgpu_check_${N}_nolean.log 2>&1
grep "mgpu_check\|MGPU_OK\|Error\|error" gpurun_out/r2_lean_mgpu_check_${N}_nolean.log | cut -c1-900 | tail -6
# one GPU: lean on / off, V and W cycles, one process
(CUDA_VISIBLE_DEVICES=0 WL=cfg3 METHODS="cg+amg+WDEPTH=0+LEAN=1,cg+amg+WDEPTH=0+LEAN=0,cg+amg+WDEPTH=2+LEAN=1,cg+amg+WDEPTH=2+LEAN=0,cg+amg+WDEPTH=3+LEAN=1" timeout 300 python tools/linsolve_probe.py 2>&1 | grep "amg" > gpurun_out/r2_lean_sweep_cfg3.log) &
(CUDA_VISIBLE_DEVICES=1 WL=cfg4 METHODS="bicgstab+amg+WDEPTH=0+LEAN=1,bicgstab+amg+WDEPTH=0+LEAN=0,bicgstab+amg+WDEPTH=2+LEAN=1" timeout 300 python tools/linsolve_probe.py 2>&1 | grep "amg" > gpurun_out/r2_lean_sweep_cfg4.log) &
wait
cat gpurun_out/r2_lean_sweep_cfg3.log gpurun_out/r2_lean_sweep_cfg4.log
run() { # tag, workload, env...
  local tag=$1; shift; local wl=$1; shift
  env "$@" timeout 300 $TR --master-port 29513 bench.py --gpus $N --workload $wl --no-cpu --no-parity --steps 5 > gpurun_out/r2_lean_${N}gpu_$tag.json 2> gpurun_out/r2_lean_${N}gpu_$tag.err
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/r2_lean_${N}gpu_$tag.json") if l.startswith("{")][-1]; n = d["newton_step"]
    print("$tag", "newton ms", round(n["ms"], 2), "iters", n["iters"], "ms/it", round(n["ms_per_iteration"], 3), "setup ms", round(n["linsolve_setup_ms"], 2), "launches", n["gpu_launches"], "res", n["resnorm"])
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/r2_lean_${N}gpu_$tag.err").read()[-1500:])
PY
}
run cfg3_v cfg3 X=1
run cfg3_w2 cfg3 VFVM_BENCH_AMG_OPTS=,,,,,2
run cfg4_w2 cfg4 VFVM_BENCH_AMG_OPTS=,,,,,2
