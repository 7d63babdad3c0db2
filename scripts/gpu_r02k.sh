#!/bin/bash
# r02k visit: exact edge test deferred through the traversal stack (tri_exact runs in the pop branch): cost per variant, ids, GPU suite incl. refit
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02k_pytest_gpu.txt
timeout 600 python scripts/debug_c3_ids.py 2>&1 | head -2 | tee gpurun_out/r02k_c3_ids.txt
{
for rep in 1 2; do
  timeout 600 python scripts/gpu_ab2.py intree C2:128 C3:64 C4:128
  for v in noexact0; do
    PTB_LIB_PATH=$PWD/build_ab/libptb200_$v.so timeout 600 python scripts/gpu_ab2.py $v C2:128 C3:64 C4:128
  done
done
} 2>&1 | grep -v "^+" | grep "pipes=" | tee gpurun_out/r02k_ab_exact_deferred.txt
ls -la gpurun_out
timeout 1200 python bench.py --steps 5 --warmup 3 2>gpurun_out/r02k_bench.err | tee gpurun_out/r02k_bench.json | cut -c1-600
