#!/bin/bash
# visit m1: parity with 1 and 2 pass pipelines, then the pipeline sweep
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/m1_pytest_pipes1.log
PTB_PIPES=2 timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/m1_pytest_pipes2.log
timeout 900 python scripts/gpu_pipes.py 2>&1 | tee gpurun_out/m1_pipes.log
