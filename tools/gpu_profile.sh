#!/bin/bash
# Run on the GPU box (under gpurun): launch list + full ncu capture of the hot kernels.
# usage: tools/gpu_profile.sh <workload> <tag>
set -u
W=${1:-C2}; TAG=${2:-r02}
mkdir -p gpurun_out
# every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv \
    --log-file gpurun_out/launches_${W}_${TAG}.csv \
    python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench_${W}.log 2>&1
# full capture of the transform's kernels (the plan build's FFT passes are skipped by name)
ncu --set full --clock-control none --import-source on \
    -k regex:'gather_kernel|estimate|select_|vote_kernel|v2_fused_kernel|v2_regroup_kernel|comb_|v3_|peel|scatter' \
    -c 24 -o gpurun_out/prof_${W}_${TAG} \
    python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_${W}.log 2>&1
ncu -i gpurun_out/prof_${W}_${TAG}.ncu-rep --page raw --csv > gpurun_out/prof_${W}_${TAG}_raw.csv 2>/dev/null
rm -f gpurun_out/prof_${W}_${TAG}.ncu-rep
