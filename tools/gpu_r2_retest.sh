#!/bin/bash
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "enable_species_after or sparse_unknown or example115 or boundary_species or devex or joule or inplace" > $O/r2_retest_parity.log 2>&1
tail -4 $O/r2_retest_parity.log
timeout 600 python -m pytest tests/test_gpu_postprocess.py -m gpu -q > $O/r2_retest_postprocess.log 2>&1
tail -4 $O/r2_retest_postprocess.log
