#!/bin/bash
# r02r visit: FHFMA node step held to 56 registers (9 resident blocks, a few spilled bytes) against the float node test; the float test at 8 blocks per SM
set -x
mkdir -p gpurun_out
{
for rep in 1 2; do
  PTB_LIB_PATH=$PWD/build_ab/libptb200_h_minb9.so timeout 600 python scripts/gpu_ab2.py h_minb9 C2:128 C3:64
  PTB_LIB_PATH=$PWD/build_ab/libptb200_f32.so timeout 600 python scripts/gpu_ab2.py f32 C2:128 C3:64
  PTB_AB_TB=8 PTB_LIB_PATH=$PWD/build_ab/libptb200_f32.so timeout 600 python scripts/gpu_ab2.py f32_tb8 C2:128 C3:64
done
} 2>&1 | grep -v "^+" | grep -E "pipes=" | tee gpurun_out/r02r_ab_node_half56.txt
