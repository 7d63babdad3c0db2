#!/bin/bash
# r02ac visit: yarn segments decided and shaded in the reference's own arithmetic (cylinder_t_rn: one rounding per operation): where
# did the YARN scene differ (scripts/debug_yarn.py), then the full GPU suite
set -x
mkdir -p gpurun_out
timeout 300 python scripts/debug_yarn.py intree 2>&1 | grep "^\[" | tee gpurun_out/r02ac_debug_yarn.txt
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -12 | tee gpurun_out/r02ac_pytest_gpu.txt
