#!/bin/bash
# r02n visit: node step of k_trace with half factors (FHFMA): GPU suite incl. the conservativeness KAT, then A/B against the float node
# test built from the same tree (f32), the half form forced to 56 registers (h_minb9) and the r02l library
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02n_pytest_gpu.txt
{
for rep in 1 2; do
  timeout 600 python scripts/gpu_ab2.py half C2:128 C3:64
  for v in f32 h_minb9 r02l; do
    PTB_LIB_PATH=$PWD/build_ab/libptb200_$v.so timeout 600 python scripts/gpu_ab2.py $v C2:128 C3:64
  done
done
} 2>&1 | grep -v "^+" | grep -E "pipes=|n_node" | tee gpurun_out/r02n_ab_node_half.txt
