#!/bin/bash
# r02ak visit: the triangle phase of k_trace in numbers (instrumented COUNT kernels, PTB_DEBUG_TRI): warp iterations, lanes per iteration and
# how few iterations a perfect packing of the same tests (32 to an iteration) would need: the upper bound of handing triangles to idle lanes
set -x
mkdir -p gpurun_out
PTB_DEBUG_TRI=1 timeout 600 python scripts/gpu_ab2.py tri C2:16 C3:16 C4:16 2>&1 | grep -E "triangle phase|closest:" | tee gpurun_out/r02ak_triangle_phase.txt
