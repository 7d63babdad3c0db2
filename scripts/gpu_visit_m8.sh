#!/bin/bash
# visit m8: commit-time alpha classification on C3 (off / on), parity suite
set -x
mkdir -p gpurun_out
PTB_ALPHA_CLASSIFY=0 timeout 300 python scripts/gpu_ab2.py alpha_classify_off C3:128 2>&1 | tee gpurun_out/m8_ab.log
timeout 300 python scripts/gpu_ab2.py alpha_classify_on C3:128 C2:256 2>&1 | tee -a gpurun_out/m8_ab.log
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/m8_pytest.log
