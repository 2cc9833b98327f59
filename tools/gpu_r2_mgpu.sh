#!/bin/bash
# N GPUs (gpurun --gpus N): multi-rank parity script + the default bench line + fused-coarse-cycle A/B
N=${1:-2}
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29512 tests/mgpu_check.py > gpurun_out/r2_mgpu_check_$N.log 2>&1
grep "mgpu_check\|MGPU_OK\|Error\|error" gpurun_out/r2_mgpu_check_$N.log | tail -12
timeout 1500 $TR --master-port 29513 bench.py --gpus $N > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -c 1500 gpurun_out/r2_bench_${N}gpu.json; tail -5 gpurun_out/r2_bench_${N}gpu.err
VFVM_AMG_NO_FUSE=1 timeout 900 $TR --master-port 29514 bench.py --gpus $N --workload cfg3 --no-cpu --no-parity --steps 5 > gpurun_out/r2_bench_${N}gpu_nofuse.json 2> gpurun_out/r2_bench_${N}gpu_nofuse.err
VFVM_NO_KRYLOV_GRAPH=1 VFVM_AMG_NO_FUSE=1 timeout 900 $TR --master-port 29515 bench.py --gpus $N --workload cfg3 --no-cpu --no-parity --steps 5 > gpurun_out/r2_bench_${N}gpu_nograph.json 2> gpurun_out/r2_bench_${N}gpu_nograph.err
python - <<PY
import json
for tag in ("", "_nofuse", "_nograph"):
    try:
        d = json.load(open(f"gpurun_out/r2_bench_${N}gpu{tag}.json"))
        n = d["newton_step"]
        print(tag or "default", "cfg3 newton ms", round(n["ms"], 2), "iters", n["iters"], "ms/it", round(n["ms_per_iteration"], 3), "asm Medges/s", round(d["value"]), "e2e", round(d["e2e"]["value"]))
        if d.get("north_star"):
            q = d["north_star"]["newton_step"]
            print("   cfg4 newton ms", round(q["ms"], 1), "iters", q["iters"], "parity", d["north_star"]["parity"])
            print("   cfg3 parity", d["parity"])
    except Exception as e:
        print(tag, "failed", e)
PY
