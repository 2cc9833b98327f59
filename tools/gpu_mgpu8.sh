#!/bin/bash
# N-GPU run (N = $1): parity script, then the cfg3 bench line with both transports
N=${1:-8}
T="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$T --master-port 29512 tests/mgpu_check.py > gpurun_out/mgpu_peer_$N.log 2>&1; grep "mgpu_check\|MGPU_OK\|Error\|error" gpurun_out/mgpu_peer_$N.log | head -5
$T --master-port 29514 bench.py --gpus $N --no-cpu > gpurun_out/bench_${N}gpu_peer.log 2>&1; tail -1 gpurun_out/bench_${N}gpu_peer.log
VFVM_NO_PEER=1 $T --master-port 29515 bench.py --gpus $N --no-cpu > gpurun_out/bench_${N}gpu_nccl.log 2>&1; tail -1 gpurun_out/bench_${N}gpu_nccl.log
