#!/bin/bash
# r02aj visit: what would a coherent bounce queue buy?  PTB_SORT_QUEUE sorts the queue of every bounce >= 1 by {octant, Morton(origin)}
# (1) or {Morton(origin), octant} (2) with a radix sort before k_trace (one pass pipeline; the sort is outside the kernel timings)
set -x
mkdir -p gpurun_out
{
PTB_SORT_QUEUE=0 timeout 600 python scripts/gpu_ab2.py unsorted C2:128 C3:64
PTB_SORT_QUEUE=1 timeout 600 python scripts/gpu_ab2.py octant_major C2:128 C3:64
PTB_SORT_QUEUE=2 timeout 600 python scripts/gpu_ab2.py origin_major C2:128 C3:64
} 2>&1 | grep -v "^+" | grep -E "pipes=1" | tee gpurun_out/r02aj_ab_sorted_queue.txt
