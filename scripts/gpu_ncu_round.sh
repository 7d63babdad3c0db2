#!/bin/bash
# ncu part of a round, summarised ON THE BOX (the raw reports exceed what gpurun copies back): launch list of bench.py, --set full of
# k_trace (C2, C3) and k_shade (C2, C4) -> profiles/<tag>_* via scripts/collect_profiles.py -> gpurun_out/<tag>/ ; raw reports dropped
set -x
TAG=${1:-r02}
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/launches_*.csv
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_C2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --also "" > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 40 -c 4 -o gpurun_out/prof_trace_C2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --also "" > gpurun_out/ncu_full_C2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 40 -c 4 -o gpurun_out/prof_trace_C3 python bench.py --steps 1 --warmup 3 --workload C3 --no-cpu-baseline --also "" > gpurun_out/ncu_full_C3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 40 -c 4 -o gpurun_out/prof_trace_C4 python bench.py --steps 1 --warmup 3 --workload C4 --no-cpu-baseline --also "" > gpurun_out/ncu_full_C4.log 2>&1
if [ "${PTB_NCU_LITE:-0}" != "1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 20 -c 2 -o gpurun_out/prof_shade_C2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --also "" > gpurun_out/ncu_full_shade.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 20 -c 2 -o gpurun_out/prof_shade_C4 python bench.py --steps 1 --warmup 3 --workload C4 --no-cpu-baseline --also "" > gpurun_out/ncu_full_shade_C4.log 2>&1
fi
python scripts/collect_profiles.py $TAG 2>&1 | tail -30
mkdir -p gpurun_out/$TAG && cp profiles/${TAG}_* profiles/roofline_traffic.json gpurun_out/$TAG/
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out gpurun_out/$TAG
