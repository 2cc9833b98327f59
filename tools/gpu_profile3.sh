#!/bin/bash
O=gpurun_out
B="python bench.py --no-cpu --no-clocks --no-newton --steps 3 --warmup 3"
ncu --set full --clock-control none --import-source on -k regex:k_assemble_rows -s 3 -c 1 -o $O/r1b_assemble_rows_cfg5 -f $B --workload cfg5 > $O/ncu_f5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_assemble_rows -s 3 -c 1 -o $O/r1b_assemble_rows_cfg2 -f $B --workload cfg2 > $O/ncu_f2.log 2>&1
ls -la $O/r1b_assemble_rows_cfg5.ncu-rep $O/r1b_assemble_rows_cfg2.ncu-rep
