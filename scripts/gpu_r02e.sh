#!/bin/bash
# r02e visit: what the exact edge test costs, variant by variant (same GPU, same visit), and the GPU suite on the current build
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02e_pytest_gpu.txt
{
for rep in 1 2; do
  timeout 600 python scripts/gpu_ab2.py intree C2:128 C3:64
  for v in noexact32 noexact0 noexact v2 v3 v4 v5 v6; do
    PTB_LIB_PATH=$PWD/build_ab/libptb200_$v.so timeout 600 python scripts/gpu_ab2.py $v C2:128 C3:64
  done
done
} 2>&1 | grep -v "^+" | grep "pipes=" | tee gpurun_out/r02e_ab_exact_variants.txt
for v in v3 v5; do PTB_LIB_PATH=$PWD/build_ab/libptb200_$v.so timeout 600 python scripts/debug_c3_ids.py 2>&1 | head -2 | tee gpurun_out/r02e_c3_ids_$v.txt; done
ls -la gpurun_out
