#!/bin/bash
# r02m visit (8 GPUs): multi-GPU tests, bench at N=8 (torchrun) and N=1 on the same box, the in-process device group
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -9; nproc
timeout 900 python -m pytest tests -x -q -m gpu -k "nccl or group or cli or resident" 2>&1 | tail -5 | tee gpurun_out/r02m_pytest_multi.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 5 --warmup 3 2>gpurun_out/r02m_bench_N8.err | tee gpurun_out/r02m_bench_N8.json | cut -c1-300
timeout 1200 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02m_bench_N1.err | tee gpurun_out/r02m_bench_N1.json | cut -c1-300
timeout 900 python scripts/gpu_group_bench.py C2:256 C5:64 2>&1 | grep "^\[group\]" | tee gpurun_out/r02m_group.txt
tail -3 gpurun_out/r02m_bench_N8.err
