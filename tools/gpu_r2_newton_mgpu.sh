#!/bin/bash
# N GPUs: Newton-step figures only (no parity probes, no CPU arm) for cfg3 and cfg4 with the default solver settings
N=${1:-8}
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for wl in cfg3 cfg4; do
  VFVM_AMG_VERBOSE=1 timeout 300 $TR --master-port 29513 bench.py --gpus $N --workload $wl --no-cpu --no-parity --steps 5 > gpurun_out/r2_newton_${N}gpu_$wl.json 2> gpurun_out/r2_newton_${N}gpu_$wl.err
  grep "vfvm amg\] rank 0" gpurun_out/r2_newton_${N}gpu_$wl.err | sort -u | cut -c1-300
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/r2_newton_${N}gpu_$wl.json") if l.startswith("{")][-1]; n = d["newton_step"]
    print("$wl", "asm Medges/s", round(d["value"]), "newton ms", round(n["ms"], 2), n["krylov"], "iters", n["iters"], "ms/it", round(n["ms_per_iteration"], 3), "setup ms", round(n["linsolve_setup_ms"], 2), "launches", n["gpu_launches"])
except Exception as e:
    print("$wl failed", e); print(open("gpurun_out/r2_newton_${N}gpu_$wl.err").read()[-1500:])
PY
done
