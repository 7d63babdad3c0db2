#!/bin/bash
# r02ab visit: Yarns (third leaf type, per-triangle t cut in the fast test, box override in the builder / refit): full GPU suite, then the
# kernel times of C2 / C3 / C4 against r02y (the t cut is a register operand now: k_trace must not move)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02ab_pytest_gpu.txt
{
for rep in 1 2; do
  timeout 600 python scripts/gpu_ab2.py r02ab C2:128 C3:64 C4:128
done
} 2>&1 | grep -v "^+" | grep -E "pipes=" | tee gpurun_out/r02ab_kernel_times.txt
