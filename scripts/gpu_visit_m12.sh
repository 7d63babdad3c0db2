#!/bin/bash
# visit m12: BVH8 emission order (breadth first vs depth first)
set -x
mkdir -p gpurun_out
PTB_BVH_DFS=0 timeout 300 python scripts/gpu_ab2.py bfs 2>&1 | tee gpurun_out/m12_ab.log
timeout 300 python scripts/gpu_ab2.py dfs 2>&1 | tee -a gpurun_out/m12_ab.log
