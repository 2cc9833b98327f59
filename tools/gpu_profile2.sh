#!/bin/bash
# launch list of one Newton step's linear solve (CG + AMG) with per-launch time and DRAM bytes
O=gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_spmv|k_smooth|k_restrict|k_prolong|k_cg|k_finalize|k_w_|k_dot|k_gal|k_blockjacobi" -s 40 -c 260 --csv --log-file $O/r1b_newton_amg_launches_cfg3.csv python bench.py --no-cpu --no-clocks --steps 3 --warmup 3 > $O/ncu_n3.log 2>&1
wc -l $O/r1b_newton_amg_launches_cfg3.csv
