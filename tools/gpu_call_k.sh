#!/bin/bash
# round 2, call K (1 GPU): v3 graph replay + 96 KB peel path: parity suite, v3 profile, C3/C1 bench lines
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/k_pytest.log 2>&1; echo "pytest rc $?" >> $O/k_pytest.log
tail -3 $O/k_pytest.log
timeout 300 python tools/v3_peel_profile.py > $O/k_v3_peel_profile.txt 2>&1
tail -2 $O/k_v3_peel_profile.txt
for wl in C3 C1; do
  timeout 600 python bench.py --workload $wl --no-extras --no-cpu-baseline > $O/k_bench_$wl.json 2> $O/k_bench_$wl.err
  cut -c1-300 $O/k_bench_$wl.json
done
