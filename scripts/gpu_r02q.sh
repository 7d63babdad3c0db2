#!/bin/bash
# r02q visit: FHFMA node step with the float fallback for steep / NaN rays, against the float node test from the same tree
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu -k "half_factors or scenes_gpu_vs_oracle or triangle_soup" 2>&1 | tail -4 | tee gpurun_out/r02q_pytest_subset.txt
{
for rep in 1 2; do
  timeout 600 python scripts/gpu_ab2.py half C2:128 C3:64
  PTB_LIB_PATH=$PWD/build_ab/libptb200_f32.so timeout 600 python scripts/gpu_ab2.py f32 C2:128 C3:64
done
} 2>&1 | grep -v "^+" | grep -E "pipes=|n_node" | tee gpurun_out/r02q_ab_node_half.txt
