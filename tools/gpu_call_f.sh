#!/bin/bash
# round 2, call F (1 GPU): parity suite on the flattened vote kernel, stage timings, plan timing
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/f_pytest.log 2>&1; echo "pytest rc $?" >> $O/f_pytest.log
tail -3 $O/f_pytest.log
: > $O/f_ab.jsonl
for wl in C1 C4 C5; do
  timeout 300 python tools/gather_ab.py $wl >> $O/f_ab.jsonl 2>> $O/f_err.txt
done
SFFTB_PLAN_TIMING=1 timeout 300 python tools/gather_ab.py C4 3 2> $O/f_plan_timing.txt | cut -c1-200
cat $O/f_plan_timing.txt
cut -c1-600 $O/f_ab.jsonl
