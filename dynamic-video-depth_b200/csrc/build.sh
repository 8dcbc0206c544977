#!/bin/bash
# Build libdvd_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libdvd_b200.so
SRCS=$(ls *.cu)
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
mkdir -p build
objs=""
pids=""
for s in $SRCS; do
  o=build/${s%.cu}.o
  objs="$objs $o"
  if [ ! -f "$o" ] || [ "$s" -nt "$o" ] || [ common.cuh -nt "$o" ] || [ ../../include/dvd_b200.h -nt "$o" ] || { [ -f tc_common.cuh ] && [ tc_common.cuh -nt "$o" ]; }; then
    ( $NVCC $FLAGS -c "$s" -o "$o" > build/${s%.cu}.log 2>&1 || { cat build/${s%.cu}.log; exit 1; } ) &
    pids="$pids $!"
  fi
done
for p in $pids; do wait $p; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o $OUT $objs -lcuda
echo "built $(realpath $OUT)"
