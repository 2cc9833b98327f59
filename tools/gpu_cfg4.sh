#!/bin/bash
# cfg4 iteration: parity of the bipolar path, bench line, ncu full capture of the row kernel
python -m pytest tests/test_gpu_parity.py -x -q -k "bipolar or sedan or two_species or 160" > gpurun_out/pytest_cfg4.log 2>&1; tail -3 gpurun_out/pytest_cfg4.log
for v in 0 1 2; do
VFVM_BIPOLAR_VARIANT=$v python bench.py --workload cfg4 --no-cpu --no-newton --no-clocks > gpurun_out/bench_cfg4_v$v.log 2>&1; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_cfg4_v$v.log").read().strip().splitlines()[-1])
print("variant $v", d["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])
PY
done
ncu --set full --clock-control none --import-source on -k regex:k_assemble_rows -s 3 -c 1 -o gpurun_out/cfg4_rows -f python bench.py --workload cfg4 --no-cpu --no-newton --no-clocks --steps 3 --warmup 3 > gpurun_out/ncu_cfg4.log 2>&1; tail -2 gpurun_out/ncu_cfg4.log
