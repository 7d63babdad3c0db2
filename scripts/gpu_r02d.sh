#!/bin/bash
# r02d visit: the reference's own triangle arithmetic for rays near an edge (tri_exact) + bit-exact camera rays: ids on C2/C3/C4 at full size,
# the parity suite, and the cost of it (A/B against the build without it)
set -x
mkdir -p gpurun_out
timeout 600 python scripts/debug_c3_ids.py 2>&1 | tail -8 | tee gpurun_out/r02d_c3_ids.txt
PTB_LIB_PATH=$PWD/build_ab/libptb200_noexact.so timeout 600 python scripts/debug_c3_ids.py 2>&1 | tail -8 | tee gpurun_out/r02d_c3_ids_noexact.txt
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02d_pytest_gpu.txt
{
for rep in 1 2; do
  timeout 600 python scripts/gpu_ab2.py exact C2:128 C3:64 C4:128
  PTB_LIB_PATH=$PWD/build_ab/libptb200_noexact.so timeout 600 python scripts/gpu_ab2.py noexact C2:128 C3:64 C4:128
done
} 2>&1 | grep -v "^+" | tee gpurun_out/r02d_ab_exact_edges.txt
ls -la gpurun_out
