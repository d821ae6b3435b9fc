#!/bin/bash
# Run on the GPU box (under gpurun): launch list + full ncu capture of the top kernels.
# usage: tools/gpu_profile.sh <workload> <tag>
set -u
W=${1:-C2}; TAG=${2:-r01}
mkdir -p gpurun_out
# every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_${W}_${TAG}.csv \
    python bench.py --workload $W --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${W}.log 2>&1
# full capture of the gather and estimate kernels (skip the plan build's launches)
ncu --set full --clock-control none --import-source on -k regex:'gather_kernel|estimate_kernel|select_kernel|fft_pass_kernel|vote_kernel' \
    -s 40 -c 10 -o gpurun_out/prof_${W}_${TAG} \
    python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${W}.log 2>&1
ls -la gpurun_out
