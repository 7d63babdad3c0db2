#!/bin/bash
# r02a visit: GPU tests on the round-2 starting state, per-bounce kernel times, A/B of the shared-memory traversal stack
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/r02a_pytest_gpu.txt
{
for rep in 1 2; do
  timeout 600 python scripts/gpu_ab2.py base C2:128 C3:64 C4:128
  for v in ss4 ss8 ss12; do
    PTB_LIB_PATH=$PWD/build_ab/libptb200_$v.so timeout 600 python scripts/gpu_ab2.py $v C2:128 C3:64 C4:128
  done
done
} 2>&1 | grep -v "^+" | tee gpurun_out/r02a_ab_smem_stack.txt
PTB_DEBUG_BOUNCES=1 PTB_PIPES=1 timeout 600 python scripts/gpu_ab2.py bounces C2:32 C3:16 > gpurun_out/r02a_bounces.txt 2>&1
ls -la gpurun_out
