#!/bin/bash
# N GPUs: replicated coarse AMG levels -- parity script, then Newton-step A/B (replication on / off, V / W cycle) on cfg3 and cfg4
N=${1:-2}
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
VFVM_AMG_VERBOSE=1 timeout 600 $TR --master-port 29512 tests/mgpu_check.py > gpurun_out/r2_repl_mgpu_check_$N.log 2>&1
grep "mgpu_check\|MGPU_OK\|Error\|error\|vfvm amg" gpurun_out/r2_repl_mgpu_check_$N.log | cut -c1-400 | tail -24
run() { # tag, env..., workload
  local tag=$1; shift; local wl=$1; shift
  env "$@" VFVM_AMG_VERBOSE=1 timeout 600 $TR --master-port 29513 bench.py --gpus $N --workload $wl --no-cpu --no-parity --steps 5 > gpurun_out/r2_repl_${N}gpu_$tag.json 2> gpurun_out/r2_repl_${N}gpu_$tag.err
  grep "vfvm amg\] rank 0" gpurun_out/r2_repl_${N}gpu_$tag.err | tail -1 | cut -c1-300
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_repl_${N}gpu_$tag.json")); n = d["newton_step"]
    print("$tag", "newton ms", round(n["ms"], 2), "iters", n["iters"], "ms/it", round(n["ms_per_iteration"], 3), "setup ms", round(n["linsolve_setup_ms"], 2), "launches", n["gpu_launches"], "res", n["resnorm"])
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/r2_repl_${N}gpu_$tag.err").read()[-1500:])
PY
}
run cfg3_repl cfg3 X=1
run cfg3_norepl cfg3 VFVM_AMG_REPL_MAX_N=0
run cfg3_repl_w2 cfg3 VFVM_BENCH_AMG_OPTS=,,,,,2
run cfg3_repl1 cfg3 VFVM_AMG_REPL_MAX_N=2000000
run cfg4_repl cfg4 X=1
run cfg4_norepl cfg4 VFVM_AMG_REPL_MAX_N=0
