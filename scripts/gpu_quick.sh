#!/bin/bash
# Short GPU visit: parity tests + bench lines (+ optional ncu of the trace kernel).  Every step is time-bounded.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
for w in ${PTB_WORKLOADS:-C2 C3}; do
  timeout 600 python bench.py --steps 2 --warmup 3 --workload $w --no-cpu-baseline 2>gpurun_out/bench_$w.err | tee gpurun_out/bench_$w.json
done
if [ "${PTB_NCU:-1}" = "1" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 42 -c 4 -o gpurun_out/prof_trace_C3 python bench.py --steps 1 --warmup 1 --workload C3 --no-cpu-baseline > gpurun_out/ncu_full_C3.log 2>&1
fi
ls -la gpurun_out | head -30
