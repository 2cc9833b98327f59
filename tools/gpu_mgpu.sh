#!/bin/bash
# 2-GPU checks: parity with peer mailboxes and with NCCL, then the cfg3 Newton step with both transports
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29512 tests/mgpu_check.py > gpurun_out/mgpu_peer.log 2>&1; grep "mgpu_check\|MGPU_OK\|Error\|error" gpurun_out/mgpu_peer.log | head -5
VFVM_NO_PEER=1 $T --master-port 29513 tests/mgpu_check.py > gpurun_out/mgpu_nccl.log 2>&1; grep "mgpu_check\|MGPU_OK\|Error\|error" gpurun_out/mgpu_nccl.log | head -5
$T --master-port 29514 bench.py --gpus 2 --no-cpu > gpurun_out/bench_2gpu_peer.log 2>&1; tail -1 gpurun_out/bench_2gpu_peer.log
VFVM_NO_PEER=1 $T --master-port 29515 bench.py --gpus 2 --no-cpu > gpurun_out/bench_2gpu_nccl.log 2>&1; tail -1 gpurun_out/bench_2gpu_nccl.log
