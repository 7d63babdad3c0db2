#!/bin/bash
# A/B of two builds of libptb200.so: usage gpu_ab.sh <other.so> [workloads...]   (the in-tree build is "new")
set -x
mkdir -p gpurun_out
OTHER=$1; shift
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_ab.log
for w in ${@:-C2 C3}; do
  for rep in 1 2; do
    PTB_LIB_PATH=$PWD/$OTHER timeout 600 python bench.py --steps 3 --warmup 3 --workload $w --no-cpu-baseline 2>>gpurun_out/ab.err | tee gpurun_out/ab_${w}_base_$rep.json | python scripts/show_bench.py /dev/stdin
    timeout 600 python bench.py --steps 3 --warmup 3 --workload $w --no-cpu-baseline 2>>gpurun_out/ab.err | tee gpurun_out/ab_${w}_new_$rep.json | python scripts/show_bench.py /dev/stdin
  done
done
