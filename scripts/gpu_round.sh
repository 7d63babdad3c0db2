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
  # launch list + full captures, summarised on the box (the raw reports exceed what gpurun copies back)
  bash scripts/gpu_ncu_round.sh ${PTB_TAG:-round}
fi
ls -la gpurun_out
