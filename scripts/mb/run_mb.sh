#!/bin/bash
# build + run the isolated contraction microbenchmark (experiments only)
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -std=c++17 -lineinfo -I../../pareben_b200/csrc $MBFLAGS -o mb_contract2 mb_contract2.cu 2>&1 | grep -v deprecated
