#!/bin/bash
# visit m4: parity on the new defaults (2 pipelines, optimal collapse, MERL float path), A/B lines, traversal knob sweep
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/m4_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/m4_smoke.log
timeout 600 python scripts/gpu_ab2.py m4 2>&1 | tee gpurun_out/m4_ab.log
timeout 600 python scripts/gpu_sweep_trace.py 2>&1 | tee gpurun_out/m4_sweep.log
