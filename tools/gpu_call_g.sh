#!/bin/bash
# round 2, call G (1 GPU): parity suite (vote aggregation, host chains, v3 large-round path), timings
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/g_pytest.log 2>&1; echo "pytest rc $?" >> $O/g_pytest.log
tail -3 $O/g_pytest.log
: > $O/g_ab.jsonl
: > $O/g_plan_timing.txt
for wl in C1 C2 C3 C4 C5; do
  echo "== $wl" >> $O/g_plan_timing.txt
  SFFTB_PLAN_TIMING=1 timeout 300 python tools/gather_ab.py $wl >> $O/g_ab.jsonl 2>> $O/g_plan_timing.txt
done
grep -v "^$" $O/g_plan_timing.txt | tail -30
timeout 300 python tools/v3_peel_profile.py > $O/g_v3_peel_profile.txt 2>&1
tail -2 $O/g_v3_peel_profile.txt
cut -c1-700 $O/g_ab.jsonl
