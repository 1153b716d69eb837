#!/bin/bash
# Build libmxb (fast) and libmxb_strict (bit-parity) for sm_100a, in-tree.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
COMMON="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -Xptxas -v"
$NVCC $COMMON -DMXB_FAST -o ../libmxb.so mxb_trace.cu 2> build_fast.log || { cat build_fast.log; exit 1; }
$NVCC $COMMON -fmad=false -o ../libmxb_strict.so mxb_trace.cu 2> build_strict.log || { cat build_strict.log; exit 1; }
grep -E "registers|spill|error|warning" build_fast.log build_strict.log | grep -v "^$" | head -40
