#!/bin/bash
# r02aa visit (8 GPUs): multi-GPU tests, bench at N=8 (torchrun; C2 main line, C3 / C4 / C5 in `also`) and a C2-only line at N=1 on the same box,
# the in-process device group on C2 / C5 (16-pixel tiles among 8 GPUs, secondary workloads warmed for a second)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -9; nproc
timeout 900 python -m pytest tests -x -q -m gpu -k "nccl or group or cli or resident" 2>&1 | tail -5 | tee gpurun_out/r02aa_pytest_multi.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 5 2>gpurun_out/r02aa_bench_N8.err | tee gpurun_out/r02aa_bench_N8.json | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --also "" 2>gpurun_out/r02aa_bench_N1.err | tee gpurun_out/r02aa_bench_N1.json | cut -c1-300
timeout 900 python scripts/gpu_group_bench.py C2:256 C5:64 2>&1 | grep "^\[group\]" | tee gpurun_out/r02aa_group.txt
tail -3 gpurun_out/r02aa_bench_N8.err
