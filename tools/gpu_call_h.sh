#!/bin/bash
# round 2, call H (1 GPU): parity suite on the current tree, cluster-barrier microbenchmark, v3 profile, stage timings
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/h_pytest.log 2>&1; echo "pytest rc $?" >> $O/h_pytest.log
tail -3 $O/h_pytest.log
tools/microbench/cluster_barrier > $O/h_cluster_barrier.jsonl 2>&1
cat $O/h_cluster_barrier.jsonl
: > $O/h_ab.jsonl
for wl in C1 C3 C4 C5; do
  timeout 300 python tools/gather_ab.py $wl >> $O/h_ab.jsonl 2>> $O/h_err.txt
done
timeout 300 python tools/v3_peel_profile.py > $O/h_v3_peel_profile.txt 2>&1
tail -2 $O/h_v3_peel_profile.txt
cut -c1-700 $O/h_ab.jsonl
