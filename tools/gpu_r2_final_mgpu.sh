#!/bin/bash
# N GPUs: multi-rank parity script + the default bench line (cfg3 + the north_star sub-run cfg4, parity objects included)
N=${1:-8}
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
VFVM_AMG_VERBOSE=1 timeout 200 $TR --master-port 29512 tests/mgpu_check.py > gpurun_out/r2_final_mgpu_check_$N.log 2>&1
grep "mgpu_check\|MGPU_OK\|Error\|error" gpurun_out/r2_final_mgpu_check_$N.log | cut -c1-700 | tail -8
VFVM_AMG_VERBOSE=1 timeout 900 $TR --master-port 29513 bench.py --gpus $N > gpurun_out/r2_final_bench_${N}gpu.json 2> gpurun_out/r2_final_bench_${N}gpu.err
grep "vfvm amg\] rank 0" gpurun_out/r2_final_bench_${N}gpu.err | sort -u | cut -c1-300
python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/r2_final_bench_${N}gpu.json") if l.startswith("{")][-1]
    n = d["newton_step"]
    print("cfg3 asm Medges/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "| newton ms", round(n["ms"], 2), n["krylov"], "iters", n["iters"], "ms/it", round(n["ms_per_iteration"], 3), "launches", n["gpu_launches"])
    print("   parity", json.dumps(d["parity"])[:900])
    q = d["north_star"]; m = q["newton_step"]
    print("cfg4 asm Medges/s", round(q["value"]), "frac", round(q["roofline"]["frac"], 3), "e2e", round(q["e2e"]["value"]), "| newton ms", round(m["ms"], 1), m["krylov"], "iters", m["iters"], "ms/it", round(m["ms_per_iteration"], 3))
    print("   parity", json.dumps(q["parity"])[:900])
except Exception as e:
    print("failed", e); print(open("gpurun_out/r2_final_bench_${N}gpu.err").read()[-2500:])
PY
