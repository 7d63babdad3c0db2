#!/bin/bash
# r02ae visit: the bench line of the final state with both arms (r02ad called a `time` binary the box does not have), wall time of the
# default run, and the yarn cloth against the oracle at several radii (where does the equal-seed image bound stop holding?)
set -x
mkdir -p gpurun_out
timeout 600 python scripts/debug_yarn_cloth.py 2>&1 | grep "^\[" | tee gpurun_out/r02ae_yarn_cloth.txt
T0=$(date +%s)
timeout 1200 python bench.py 2>gpurun_out/r02ae_bench.err | tee gpurun_out/r02ae_bench.json | cut -c1-300
echo "default bench.py run: $(( $(date +%s) - T0 )) s wall" | tee gpurun_out/r02ae_bench_time.txt
timeout 600 python bench.py --impl reference 2>>gpurun_out/r02ae_bench.err | tee gpurun_out/r02ae_bench_reference.json | cut -c1-300
