#!/bin/bash
# visit m2: source-level ncu captures of k_shade (Phong on C2, MERL on C4) and of k_trace on C2
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 20 -c 2 -o gpurun_out/m2_shade_C2 python bench.py --steps 1 --warmup 1 --workload C2 --no-cpu-baseline > gpurun_out/m2_ncu_shade_C2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 20 -c 2 -o gpurun_out/m2_shade_C4 python bench.py --steps 1 --warmup 1 --workload C4 --no-cpu-baseline > gpurun_out/m2_ncu_shade_C4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 40 -c 2 -o gpurun_out/m2_trace_C2 python bench.py --steps 1 --warmup 1 --workload C2 --no-cpu-baseline > gpurun_out/m2_ncu_trace_C2.log 2>&1
ls -la gpurun_out
