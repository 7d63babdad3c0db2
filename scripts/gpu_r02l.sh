#!/bin/bash
# r02l visit: ncu round on the final kernels (launch list, --set full of k_trace C2/C3/C4 and k_shade C2/C4 -> profiles/r02l_*, roofline_traffic.json), then the bench line
set -x
mkdir -p gpurun_out
bash scripts/gpu_ncu_round.sh r02l 2>&1 | tail -60
timeout 1200 python bench.py --steps 5 --warmup 3 2>gpurun_out/r02l_bench.err | tee gpurun_out/r02l_bench.json | cut -c1-400
