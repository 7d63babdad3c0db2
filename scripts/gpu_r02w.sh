#!/bin/bash
# r02w visit: A/B of small kernel variants on the r02v build: k_shade's two queue atomics issued together (in-tree), the node step's hit mask
# accumulated on the fma pipe, k_shade blocks of 256 / 512 threads, k_shade at 7 resident blocks (72 registers)
set -x
mkdir -p gpurun_out
{
for rep in 1 2; do
  timeout 600 python scripts/gpu_ab2.py intree C2:128 C3:64 C4:128
  for v in maskfma sb256 sb512; do
    PTB_LIB_PATH=$PWD/build_ab/libptb200_$v.so timeout 600 python scripts/gpu_ab2.py $v C2:128 C3:64 C4:128
  done
  PTB_SHADE_MINB=7 PTB_SHADE_MINB_MERL=7 timeout 600 python scripts/gpu_ab2.py minb7 C2:128 C4:128
done
} 2>&1 | grep -v "^+" | grep -E "pipes=" | tee gpurun_out/r02w_ab_variants.txt
