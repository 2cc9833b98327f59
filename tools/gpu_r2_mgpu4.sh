#!/bin/bash
N=${1:-4}
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29512 tests/mgpu_check.py > gpurun_out/r2_mgpu_check_$N.log 2>&1
grep "mgpu_check\|MGPU_OK" gpurun_out/r2_mgpu_check_$N.log | tail -8 | cut -c1-600
timeout 600 $TR --master-port 29514 bench.py --gpus $N --workload cfg3 --no-cpu --steps 5 > gpurun_out/r2_bench_${N}gpu_fp32.json 2> gpurun_out/r2_bench_${N}gpu_fp32.err
VFVM_AMG_FP32=0 timeout 600 $TR --master-port 29515 bench.py --gpus $N --workload cfg3 --no-cpu --no-parity --steps 5 > gpurun_out/r2_bench_${N}gpu_fp64.json 2> gpurun_out/r2_bench_${N}gpu_fp64.err
python - <<PY
import json
for tag in ("_fp32", "_fp64"):
    try:
        txt = open(f"gpurun_out/r2_bench_${N}gpu{tag}.json").read()
        d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
        n = d["newton_step"]
        print(tag, "cfg3 newton ms", round(n["ms"], 2), "iters", n["iters"], "ms/it", round(n["ms_per_iteration"], 3), "launches", n["gpu_launches"], "asm Medges/s", round(d["value"]))
        if d.get("parity"):
            print("   parity:", d["parity"]["assembly"]["ok"], d["parity"]["newton"])
    except Exception as e:
        print(tag, "failed", repr(e))
PY
tail -3 gpurun_out/r2_bench_${N}gpu_fp32.err
