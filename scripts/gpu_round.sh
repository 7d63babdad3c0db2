#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench lines, ncu launch list + full capture of the top kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py --steps 3 --warmup 3 2>gpurun_out/bench_C2.err | tee gpurun_out/bench_C2.json
python bench.py --steps 2 --warmup 3 --workload C3 --no-cpu-baseline 2>gpurun_out/bench_C3.err | tee gpurun_out/bench_C3.json
python bench.py --steps 2 --warmup 3 --workload C4 --no-cpu-baseline 2>gpurun_out/bench_C4.err | tee gpurun_out/bench_C4.json
python bench.py --steps 3 --warmup 3 --workload C1 --no-cpu-baseline 2>gpurun_out/bench_C1.err | tee gpurun_out/bench_C1.json
if [ "${PTB_NCU:-1}" = "1" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 330 --csv --log-file gpurun_out/launches_C2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 21 -c 3 -o gpurun_out/prof_extend_C2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_extend -s 21 -c 2 -o gpurun_out/prof_extend_C3 python bench.py --steps 1 --warmup 1 --workload C3 --no-cpu-baseline > gpurun_out/ncu_full_C3.log 2>&1
fi
ls -la gpurun_out
