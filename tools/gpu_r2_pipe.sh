#!/bin/bash
# N GPUs: pipelined host-vector assembly per rank -- parity script (bitwise against the plain path), then the e2e figures of cfg3 / cfg4 with and without
N=${1:-2}
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29512 tests/mgpu_check.py > gpurun_out/r2_pipe_mgpu_check_$N.log 2>&1
grep "mgpu_check\|MGPU_OK\|Error\|error" gpurun_out/r2_pipe_mgpu_check_$N.log | cut -c1-500 | tail -8
run() { # tag, workload, env...
  local tag=$1; shift; local wl=$1; shift
  env "$@" timeout 300 $TR --master-port 29513 bench.py --gpus $N --workload $wl --no-cpu --no-parity --no-newton --steps 10 > gpurun_out/r2_pipe_${N}gpu_$tag.json 2> gpurun_out/r2_pipe_${N}gpu_$tag.err
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/r2_pipe_${N}gpu_$tag.json") if l.startswith("{")][-1]
    print("$tag", "asm Medges/s", round(d["value"]), "e2e Medges/s", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 3))
except Exception as e:
    print("$tag failed", e); print(open("gpurun_out/r2_pipe_${N}gpu_$tag.err").read()[-1500:])
PY
}
run cfg3_pipe cfg3 X=1
run cfg3_plain cfg3 VFVM_NO_PIPELINE=1
run cfg4_pipe cfg4 X=1
run cfg4_plain cfg4 VFVM_NO_PIPELINE=1
