#!/bin/bash
# Builds libvfvmb200.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ $ARCH"
OUT=../libvfvmb200.so
mkdir -p build
pids=()
compile() { # src extra-flags
  local src=$1; shift
  if [ ! -f build/${src%.cu}.o ] || [ $src -nt build/${src%.cu}.o ] || [ vfvm_internal.h -nt build/${src%.cu}.o ] || [ physics.cuh -nt build/${src%.cu}.o ] || [ dual.cuh -nt build/${src%.cu}.o ] || [ peer.cuh -nt build/${src%.cu}.o ] || [ ../../include/vfvm_b200.h -nt build/${src%.cu}.o ]; then
    $NVCC $COMMON "$@" -c $src -o build/${src%.cu}.o &
    pids+=($!)
  fi
}
# geometry: no FMA contraction so the per-simplex form factors round like a plain fp64 CPU evaluation
compile geometry.cu -fmad=false
compile pattern.cu
compile assemble.cu ${VFVM_PTXAS_V:+-Xptxas -v}
compile linsolve.cu
compile ilu0.cu
compile comm.cu
compile amg.cu
compile postprocess.cu
compile api.cu
for p in "${pids[@]}"; do wait $p; done
$NVCC $ARCH -shared -cudart static -o $OUT build/geometry.o build/pattern.o build/assemble.o build/linsolve.o build/ilu0.o build/comm.o build/amg.o build/postprocess.o build/api.o -ldl
echo "built $(realpath $OUT)"
