#!/bin/bash
# ncu evidence for profiles/: launch lists and --set full captures of the dominant kernels (1 GPU; numbers under ncu are never bench values)
O=gpurun_out
B="python bench.py --no-cpu --no-clocks --steps 3 --warmup 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r1b_launches_cfg3.csv $B > $O/ncu_l3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r1b_launches_cfg4.csv $B --workload cfg4 --no-newton > $O/ncu_l4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_assemble_rows -s 3 -c 1 -o $O/r1b_assemble_rows_cfg3 -f $B --no-newton > $O/ncu_f3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_assemble_rows -s 3 -c 1 -o $O/r1b_assemble_rows_bipolar_cfg4 -f $B --workload cfg4 --no-newton > $O/ncu_f4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_spmv|k_smooth|k_restrict|k_prolong|k_gal_off|k_gal_diag|k_cg_xr" -s 30 -c 12 -o $O/r1b_newton_amg_cfg3 -f $B > $O/ncu_fn.log 2>&1
ls -la $O/*.ncu-rep
