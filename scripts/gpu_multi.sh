#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): sharded == single check, then bench lines at N and at 1 for C2 and C5.
set -x
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR scripts/multi_gpu_check.py 2>gpurun_out/multi_check_$N.err | tee gpurun_out/multi_check_$N.json
for w in ${PTB_WORKLOADS:-C2 C5}; do
  timeout 1200 $TR bench.py --gpus $N --steps 2 --warmup 3 --workload $w 2>gpurun_out/bench_${w}_N$N.err | tee gpurun_out/bench_${w}_N$N.json
done
if [ "${PTB_SINGLE:-1}" = "1" ]; then
  timeout 1200 python bench.py --gpus 1 --steps 2 --warmup 3 --workload C5 --no-cpu-baseline 2>gpurun_out/bench_C5_N1.err | tee gpurun_out/bench_C5_N1.json
fi
for f in gpurun_out/*.err; do tail -n 3 "$f"; done
