#!/bin/bash
# r02b visit: new parity tests (C5 vs oracle at size, converged at size, stack limit, group / resident renders), the reworked bench line
# (also: C3 / C4 / C5, reference arm at fixed spp), smem-only stack variant
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
nproc
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -25 | tee gpurun_out/r02b_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r02b_smoke.log
timeout 1200 python bench.py --steps 5 --warmup 3 2>gpurun_out/r02b_bench.err | tee gpurun_out/r02b_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/r02b_bench_ref.err | tee gpurun_out/r02b_bench_ref.json
PTB_LIB_PATH=$PWD/build_ab/libptb200_sso20.so timeout 600 python scripts/gpu_ab2.py sso20 C2:128 C3:64 2>&1 | grep -v "^+" | tee gpurun_out/r02b_ab_smem_only.txt
tail -5 gpurun_out/r02b_bench.err
ls -la gpurun_out
