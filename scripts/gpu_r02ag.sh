#!/bin/bash
# r02ag visit: cache-policy A/B on the final kernels, same box: k_trace's triangle loads without L1 allocation / evict-first, node loads
# evict-last, the path pool's loads and stores with the streaming hint (PTB_LD_TRI, PTB_LD_NODE, PTB_POOL_STREAM); images must not move
set -x
mkdir -p gpurun_out
{
timeout 600 python scripts/gpu_ab2.py intree C2:128 C3:64 C4:128
for v in tri1 tri2n node1 pool pooltri; do
  PTB_LIB_PATH=$PWD/build_ab/libptb200_$v.so timeout 600 python scripts/gpu_ab2.py $v C2:128 C3:64 C4:128
done
timeout 600 python scripts/gpu_ab2.py intree C2:128 C3:64 C4:128
} 2>&1 | grep -v "^+" | grep -E "pipes=" | tee gpurun_out/r02ag_ab_cache_hints.txt
