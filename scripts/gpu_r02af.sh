#!/bin/bash
# r02af visit (2 GPUs): the whole GPU suite incl. the NCCL / device-group cases on the final state, then bench.py at N=2 under torchrun
# (the brief secondary workloads now run one untimed full step first) and at N=1 on the same box
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -3
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02af_pytest_gpu_2gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 2 --steps 3 --warmup 3 2>gpurun_out/r02af_bench_N2.err | tee gpurun_out/r02af_bench_N2.json | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02af_bench_N1.err | tee gpurun_out/r02af_bench_N1.json | cut -c1-300
tail -3 gpurun_out/r02af_bench_N2.err
