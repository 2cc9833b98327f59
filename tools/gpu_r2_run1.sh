#!/bin/bash
# round 2, first GPU call (1 GPU): device Newton goldens, GPU test suite, default bench line, launch list of one Newton step
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_run1_gpu.txt
nproc >> gpurun_out/r2_run1_gpu.txt; free -g >> gpurun_out/r2_run1_gpu.txt
timeout 600 python tools/make_newton_golden.py --workload cfg4 --source device --out gpurun_out/newton_samples_cfg4_193.json > gpurun_out/r2_golden_cfg4.log 2>&1
cp gpurun_out/newton_samples_cfg4_193.json tests/golden/ 2>/dev/null
timeout 600 python tools/make_newton_golden.py --workload cfg3 --source device --out gpurun_out/newton_samples_cfg3_193_device.json > gpurun_out/r2_golden_cfg3.log 2>&1
[ -f tests/golden/newton_samples_cfg3_193.json ] || cp gpurun_out/newton_samples_cfg3_193_device.json tests/golden/newton_samples_cfg3_193.json
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_run1_pytest.log 2>&1
tail -5 gpurun_out/r2_run1_pytest.log
timeout 900 python bench.py > gpurun_out/r2_run1_bench.json 2> gpurun_out/r2_run1_bench.err
tail -c 3000 gpurun_out/r2_run1_bench.json
tail -5 gpurun_out/r2_run1_bench.err
