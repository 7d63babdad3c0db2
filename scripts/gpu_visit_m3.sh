#!/bin/bash
# visit m3: optimal-SAH collapse vs greedy collapse, identity-transform fast path, pipelines; then parity on the new defaults
set -x
mkdir -p gpurun_out
PTB_BVH_COLLAPSE_DP=0 timeout 600 python scripts/gpu_ab2.py greedy 2>&1 | tee gpurun_out/m3_ab.log
timeout 600 python scripts/gpu_ab2.py dp 2>&1 | tee -a gpurun_out/m3_ab.log
PTB_BVH_CPRIM=0.5 timeout 600 python scripts/gpu_ab2.py dp_cprim0.5 C2:256 C3:128 2>&1 | tee -a gpurun_out/m3_ab.log
PTB_BVH_CPRIM=0.15 timeout 600 python scripts/gpu_ab2.py dp_cprim0.15 C2:256 C3:128 2>&1 | tee -a gpurun_out/m3_ab.log
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/m3_pytest.log
