#!/bin/bash
# round 2, call M (1 GPU): the bench line of every BASELINE config (with the CPU baseline leg), then ncu evidence
set -u
mkdir -p gpurun_out
O=gpurun_out
for wl in C2 C1 C3 C4 C5 C2b; do
  timeout 900 python bench.py --workload $wl --no-extras > $O/m_bench_$wl.json 2> $O/m_bench_$wl.err
  cut -c1-200 $O/m_bench_$wl.json
done
for wl in C2 C4 C5 C1 C3; do
  bash tools/gpu_profile.sh $wl r02
done
ls -la $O | tail -30
