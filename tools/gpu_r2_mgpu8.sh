#!/bin/bash
# N GPUs: multi-rank parity script, default bench line (cfg3 + north_star cfg4, parity at full size), W-cycle A/B on cfg3
N=${1:-8}
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29512 tests/mgpu_check.py > gpurun_out/r2_mgpu_check_$N.log 2>&1
grep "mgpu_check\|MGPU_OK\|Error\|error" gpurun_out/r2_mgpu_check_$N.log | tail -8
timeout 900 $TR --master-port 29513 bench.py --gpus $N > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
VFVM_BENCH_AMG_OPTS=",,,,,2" timeout 600 $TR --master-port 29514 bench.py --gpus $N --workload cfg3 --no-cpu --no-parity --steps 5 > gpurun_out/r2_bench_${N}gpu_w2.json 2> gpurun_out/r2_bench_${N}gpu_w2.err
VFVM_AMG_FUSE=1 timeout 600 $TR --master-port 29515 bench.py --gpus $N --workload cfg3 --no-cpu --no-parity --steps 5 > gpurun_out/r2_bench_${N}gpu_fuse.json 2> gpurun_out/r2_bench_${N}gpu_fuse.err
python - <<PY
import json
for tag in ("", "_w2", "_fuse"):
    try:
        txt = open(f"gpurun_out/r2_bench_${N}gpu{tag}.json").read()
        d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
        n = d["newton_step"]
        print(tag or "default", "cfg3 newton ms", round(n["ms"], 2), "iters", n["iters"], "ms/it", round(n["ms_per_iteration"], 3), "launches", n["gpu_launches"], "asm Medges/s", round(d["value"]), "e2e", round(d["e2e"]["value"]))
        if d.get("parity"):
            print("   cfg3 parity ok:", d["parity"]["assembly"]["ok"], d["parity"]["newton"].get("max_abs_diff"))
        if d.get("north_star"):
            q = d["north_star"]; nn = q["newton_step"]
            print("   cfg4 asm", round(q["value"]), "frac", round(q["roofline"]["frac"], 3), "newton ms", round(nn["ms"], 1), "iters", nn["iters"], "ms/it", round(nn["ms_per_iteration"], 2), "parity ok:", q["parity"]["assembly"]["ok"], q["parity"]["newton"].get("max_abs_diff"))
    except Exception as e:
        print(tag, "failed", repr(e))
PY
tail -3 gpurun_out/r2_bench_${N}gpu.err
