#!/bin/bash
# Builds libpgo_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NCCL_INC=${NCCL_INC:-/usr/include}
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
     -Xcompiler -fPIC -Xptxas -v --shared -I"$NCCL_INC" \
     -o libpgo_b200.so pgo_b200.cu -lnccl -lcudart 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning: v|Used|spill" build.log | grep -v "^$" | head -80
