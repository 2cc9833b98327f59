#!/bin/bash
# one GPU: register / occupancy variants of the bipolar row kernel on cfg4 at 193^3 (VFVM_BIPOLAR_VARIANT)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for V in 0 1 2 3 4 0; do
  VFVM_BIPOLAR_VARIANT=$V timeout 200 python bench.py --workload cfg4 --no-cpu --no-parity --no-newton --no-clocks --steps 20 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('variant $V', 'step ms', round(d['ms_per_step'], 4), 'kernel ms', round(d['roofline']['kernel_ms'], 4), 'frac', round(d['roofline']['frac'], 4), 'Medges/s', round(d['value']))"
done | tee gpurun_out/r2_bipolar_variants.txt
