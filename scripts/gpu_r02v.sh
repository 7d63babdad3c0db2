#!/bin/bash
# r02v visit: in-tree library = -prec-div=false -prec-sqrt=false, exp2f(y log2f x) for the Phong powers, sincospif: GPU suite, kernel times, bench line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 | tee gpurun_out/r02v_pytest_gpu.txt
{
for rep in 1 2; do
  timeout 600 python scripts/gpu_ab2.py r02v C2:128 C3:64 C4:128
done
} 2>&1 | grep -v "^+" | grep -E "pipes=" | tee gpurun_out/r02v_ab.txt
timeout 1200 python bench.py --steps 5 --warmup 3 2>gpurun_out/r02v_bench.err | tee gpurun_out/r02v_bench.json | cut -c1-200
