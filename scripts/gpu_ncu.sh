#!/bin/bash
# ncu --set full of selected kernels on one workload.  usage: gpu_ncu.sh <workload> <kernel-regex> <skip> <count> <outname>
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o gpurun_out/$5 python bench.py --steps 1 --warmup 1 --workload $1 --no-cpu-baseline > gpurun_out/ncu_$5.log 2>&1
ls -la gpurun_out/$5.ncu-rep
