#!/bin/bash
# r02u visit: the library compiled with -prec-div=false -prec-sqrt=false (no IEEE division / square-root sequences outside the explicitly rounded code): GPU suite + A/B
set -x
mkdir -p gpurun_out
PTB_LIB_PATH=$PWD/build_ab/libptb200_fastdiv.so timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r02u_pytest_gpu_fastdiv.txt
{
for rep in 1 2; do
  timeout 600 python scripts/gpu_ab2.py base C2:128 C3:64 C4:128
  PTB_LIB_PATH=$PWD/build_ab/libptb200_fastdiv.so timeout 600 python scripts/gpu_ab2.py fastdiv C2:128 C3:64 C4:128
done
} 2>&1 | grep -v "^+" | grep -E "pipes=" | tee gpurun_out/r02u_ab_fastdiv.txt
