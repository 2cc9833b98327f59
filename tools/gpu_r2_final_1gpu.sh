#!/bin/bash
# one GPU: the driver's round-end sequence (pytest -m gpu, smoke, bench) + the ncu evidence of the same commands
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_final_pytest.log 2>&1
tail -6 $O/r2_final_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --no-cpu > $O/r2_final_bench_1gpu.json 2> $O/r2_final_bench_1gpu.err
python - <<PY
import json
try:
    d = [json.loads(l) for l in open("$O/r2_final_bench_1gpu.json") if l.startswith("{")][-1]
    n = d["newton_step"]
    print("cfg3 asm Medges/s", round(d["value"]), "frac", round(d["roofline"]["frac"], 3), "e2e", round(d["e2e"]["value"]), "| newton ms", round(n["ms"], 2), n["krylov"], "iters", n["iters"], "ms/it", round(n["ms_per_iteration"], 3), "launches", n["gpu_launches"], "clocks", d["clocks"])
    print("   parity", json.dumps(d["parity"])[:700])
    q = d["north_star"]; m = q["newton_step"]
    print("cfg4 asm Medges/s", round(q["value"]), "frac", round(q["roofline"]["frac"], 3), "e2e", round(q["e2e"]["value"]), "| newton ms", round(m["ms"], 1), m["krylov"], "iters", m["iters"], "ms/it", round(m["ms_per_iteration"], 3))
    print("   parity", json.dumps(q["parity"])[:700])
except Exception as e:
    print("failed", e); print(open("$O/r2_final_bench_1gpu.err").read()[-2500:])
PY
# launch list of the bench command (assembly steps) and of a Newton-step solve (CG + AMG, W-cycle) -- per-launch times under ncu are cold-cache and serialised
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_bench_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-north-star --no-newton --no-clocks > $O/r2_ncu_bench.log 2>&1
python tools/launch_summary.py $O/r2_launches_bench_cfg3.csv 2>&1 | head -14 | tee $O/r2_launches_bench_cfg3_summary.txt
WL=cfg3 METHODS="cg+amg+WDEPTH=2" MAXIT=6 timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file $O/r2_launches_newton_cfg3_w2.csv python tools/linsolve_probe.py > $O/r2_ncu_newton.log 2>&1
python tools/launch_summary.py $O/r2_launches_newton_cfg3_w2.csv 2>&1 | head -24 | tee $O/r2_launches_newton_cfg3_w2_summary.txt
# one full capture of the headline kernel for the round
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_assemble_rows -s 3 -c 1 -o $O/r2_assemble_rows_cfg3 -f python bench.py --no-cpu --no-clocks --no-newton --no-parity --no-north-star --steps 3 --warmup 3 > $O/r2_ncu_full.log 2>&1
ls -la $O/r2_assemble_rows_cfg3.ncu-rep
