#!/bin/bash
# r02p visit: why the FHFMA node step was slow in r02n: conversions with directed rounding (F2F.F16.F32.RM/.RP per ray) against round-to-nearest + fix-up
set -x
mkdir -p gpurun_out
./scripts/ubench/pipe_rates 2>&1 | cut -c1-90 | tee gpurun_out/r02p_pipe_rates.txt
timeout 300 python -m pytest tests -x -q -m gpu -k "half_factors or scenes_gpu_vs_oracle or triangle_soup" 2>&1 | tail -4 | tee gpurun_out/r02p_pytest_subset.txt
{
for rep in 1 2; do
  timeout 600 python scripts/gpu_ab2.py half C2:128 C3:64
  for v in f32 h_cvt; do
    PTB_LIB_PATH=$PWD/build_ab/libptb200_$v.so timeout 600 python scripts/gpu_ab2.py $v C2:128 C3:64
  done
done
} 2>&1 | grep -v "^+" | grep -E "pipes=" | tee gpurun_out/r02p_ab_node_half.txt
# source-level stall samples of the half node step, whatever the outcome (read with scripts/sass_by_line.py / sass_mix.py)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 0 -c 2 -o gpurun_out/r02p_k_trace_half python scripts/gpu_ab2.py ncu C2:4 > gpurun_out/r02p_ncu.log 2>&1
ls -la gpurun_out/r02p_k_trace_half.ncu-rep
