#!/bin/bash
# r02o visit: issue rates of the candidate instructions of a node test (scripts/ubench/pipe_rates.cu); float node test with one PRMT per plane
set -x
mkdir -p gpurun_out
./scripts/ubench/pipe_rates 2>&1 | tee gpurun_out/r02o_pipe_rates.txt
{
for rep in 1 2; do
  for v in f32 f32_prmt32; do
    PTB_LIB_PATH=$PWD/build_ab/libptb200_$v.so timeout 600 python scripts/gpu_ab2.py $v C2:128 C3:64
  done
done
} 2>&1 | grep -v "^+" | grep -E "pipes=" | tee gpurun_out/r02o_ab_prmt32.txt
