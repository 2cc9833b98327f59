#!/bin/bash
run() { # label env... workload
  python bench.py --no-cpu --no-newton --no-clocks --workload $1 > gpurun_out/tune.log 2>&1
  python - <<PY
import json
d=json.loads(open("gpurun_out/tune.log").read().strip().splitlines()[-1])
print("$1 $2", round(d["value"]), "Medges/s  kernel_ms", round(d["roofline"]["kernel_ms"],4), "frac", round(d["roofline"]["frac"],3))
PY
}
run cfg2 batch6
VFVM_SEP_BATCH8=1 run cfg2 batch8
run cfg5 ch5
VFVM_SEP_CH2=1 run cfg5 ch2
python -m pytest tests/test_gpu_parity.py -x -q -k "207 or many_species" 2>&1 | tail -2
