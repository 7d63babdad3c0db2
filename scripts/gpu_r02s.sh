#!/bin/bash
# r02s visit: GPU suite on the new node quantisation (exponents for the padded box, half grid bytes), shard balance at finer tiles, the bench line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02s_pytest_gpu.txt
PTB_TILES=32,16,8 timeout 600 python scripts/gpu_shard_balance.py C2 8 2>&1 | grep "per-shard" | tee gpurun_out/r02s_shard_balance.txt
timeout 1200 python bench.py --steps 5 --warmup 3 2>gpurun_out/r02s_bench.err | tee gpurun_out/r02s_bench.json | cut -c1-300
