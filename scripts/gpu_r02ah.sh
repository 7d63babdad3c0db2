#!/bin/bash
# r02ah visit: compute-sanitizer (memcheck) over the small yarn / point-set / refit cases of the GPU suite, then the whole GPU suite on the
# build with the pool accessors (flags off: same SASS for k_trace)
set -x
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -q -x -k "YARN or yarn_seen or refit_moves or PTS or CYL" 2>&1 | tail -15 | tee gpurun_out/r02ah_memcheck.txt
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02ah_pytest_gpu.txt
