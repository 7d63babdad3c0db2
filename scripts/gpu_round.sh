#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench lines, ncu launch list + full capture of the dominant kernel.
# Every step is time-bounded.  Outputs land in gpurun_out/ (scratch); scripts/collect_profiles.py copies what is judged into profiles/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_C2.err | tee gpurun_out/bench_C2.json
for w in ${PTB_WORKLOADS:-C3 C4 C1}; do
  timeout 900 python bench.py --steps 2 --warmup 3 --workload $w --no-cpu-baseline 2>gpurun_out/bench_$w.err | tee gpurun_out/bench_$w.json
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/bench_ref_C2.err | tee gpurun_out/bench_ref_C2.json
if [ "${PTB_NCU:-1}" = "1" ]; then
  # launch list of the bench command (durations only; cold-cache, serialised: compare SHARES, not absolutes)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_C2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
  # full capture of the closest-hit trace kernel (bounce 0 and 1 of one pass) and of k_shade
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 40 -c 4 -o gpurun_out/prof_trace_C2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_C2.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 40 -c 4 -o gpurun_out/prof_trace_C3 python bench.py --steps 1 --warmup 3 --workload C3 --no-cpu-baseline > gpurun_out/ncu_full_C3.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 20 -c 2 -o gpurun_out/prof_shade_C2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_shade.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 20 -c 2 -o gpurun_out/prof_shade_C4 python bench.py --steps 1 --warmup 3 --workload C4 --no-cpu-baseline > gpurun_out/ncu_full_shade_C4.log 2>&1
fi
ls -la gpurun_out
