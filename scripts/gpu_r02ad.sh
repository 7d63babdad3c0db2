#!/bin/bash
# r02ad visit: the final state of round 2: smoke(), the new yarn cases, the bench line with both arms (wall time of the default run noted),
# then the ncu round (launch list, --set full summaries of k_trace C2/C3/C4 and k_shade C2/C4 -> profiles/r02ad_*, roofline_traffic.json)
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r02ad_smoke.txt
timeout 600 python -m pytest tests -q -m gpu -k "yarn or YARN or errors" 2>&1 | tail -4 | tee gpurun_out/r02ad_pytest_yarn.txt
/usr/bin/time -v -o gpurun_out/r02ad_bench_time.txt timeout 1200 python bench.py 2>gpurun_out/r02ad_bench.err | tee gpurun_out/r02ad_bench.json | cut -c1-300
grep -E "Elapsed|Maximum resident" gpurun_out/r02ad_bench_time.txt
timeout 600 python bench.py --impl reference 2>>gpurun_out/r02ad_bench.err | tee gpurun_out/r02ad_bench_reference.json | cut -c1-300
bash scripts/gpu_ncu_round.sh r02ad 2>&1 | tail -25
