#!/bin/bash
# N-GPU checks (default 2): parity with peer mailboxes and with NCCL, then the cfg3 bench line with both transports
N=${1:-2}
export VFVM_AMG_VERBOSE=1
T="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$T --master-port 29512 tests/mgpu_check.py > gpurun_out/mgpu_peer.log 2>&1; grep "mgpu_check\|MGPU_OK\|Error\|error\|vfvm amg" gpurun_out/mgpu_peer.log | head -8
VFVM_NO_PEER=1 $T --master-port 29513 tests/mgpu_check.py > gpurun_out/mgpu_nccl.log 2>&1; grep "mgpu_check\|MGPU_OK\|Error\|error" gpurun_out/mgpu_nccl.log | head -5
$T --master-port 29514 bench.py --gpus $N --no-cpu > gpurun_out/bench_${N}gpu_peer.log 2>&1; grep "vfvm amg" gpurun_out/bench_${N}gpu_peer.log | head -3; tail -1 gpurun_out/bench_${N}gpu_peer.log
VFVM_NO_PEER=1 $T --master-port 29515 bench.py --gpus $N --no-cpu > gpurun_out/bench_${N}gpu_nccl.log 2>&1; tail -1 gpurun_out/bench_${N}gpu_nccl.log
