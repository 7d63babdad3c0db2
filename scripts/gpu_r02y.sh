#!/bin/bash
# r02y visit: in-tree = r02w + path_sample fast path + k_trace without the stack guards / the redundant quorum ballot: GPU suite, kernel times
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02y_pytest_gpu.txt
{
for rep in 1 2; do
  timeout 600 python scripts/gpu_ab2.py r02y C2:128 C3:64 C4:128
done
} 2>&1 | grep -v "^+" | grep -E "pipes=" | tee gpurun_out/r02y_kernel_times.txt
