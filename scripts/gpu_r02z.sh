#!/bin/bash
# r02z visit: ncu round on the final kernels of round 2 (launch list, --set full summaries of k_trace C2/C3/C4 and k_shade C2/C4 -> profiles/r02z_*, roofline_traffic.json), then the bench line with both arms
set -x
mkdir -p gpurun_out
bash scripts/gpu_ncu_round.sh r02z 2>&1 | tail -40
timeout 1200 python bench.py --steps 5 --warmup 3 2>gpurun_out/r02z_bench.err | tee gpurun_out/r02z_bench.json | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/r02z_bench.err | tee gpurun_out/r02z_bench_reference.json | cut -c1-300
