#!/bin/bash
# visit m7: traversal with next-node prefetch (build_ab/lib_prefetch.so) vs the in-tree build; grid-stride k_shade
set -x
mkdir -p gpurun_out
timeout 300 python scripts/gpu_ab2.py base 2>&1 | tee gpurun_out/m7_ab.log
PTB_LIB_PATH=$PWD/build_ab/lib_prefetch.so timeout 300 python scripts/gpu_ab2.py prefetch 2>&1 | tee -a gpurun_out/m7_ab.log
for g in 2368 4736 9472; do
  PTB_SHADE_GRID=$g timeout 300 python scripts/gpu_ab2.py shade_grid$g C2:256 C4:256 2>&1 | tee -a gpurun_out/m7_ab.log
done
PTB_LIB_PATH=$PWD/build_ab/lib_prefetch.so timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/m7_pytest_prefetch.log
