#!/bin/bash
N=${1:-8}
T="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
VFVM_AMG_WDEPTH=2 $T --master-port 29514 bench.py --gpus $N --no-cpu --steps 10 > gpurun_out/bench_${N}gpu_peer_w2.log 2>&1; tail -1 gpurun_out/bench_${N}gpu_peer_w2.log
