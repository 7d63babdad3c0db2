#!/bin/bash
# r02c visit (2 GPUs): the whole GPU suite where the NCCL / group tests run, bench at N=2 (torchrun) and N=1 on the same box
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv
nproc
timeout 1500 python -m pytest tests -x -q -m gpu --durations=6 2>&1 | tail -22 | tee gpurun_out/r02c_pytest_gpu_2gpus.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/r02c_bench_N2.err | tee gpurun_out/r02c_bench_N2.json
timeout 1200 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02c_bench_N1.err | tee gpurun_out/r02c_bench_N1.json
tail -3 gpurun_out/r02c_bench_N2.err gpurun_out/r02c_bench_N1.err
ls -la gpurun_out
