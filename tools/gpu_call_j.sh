#!/bin/bash
# round 2, call J (1 GPU): v3 parity + timing after the hash-grouped shared-memory peel path; full suite
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/j_pytest.log 2>&1; echo "pytest rc $?" >> $O/j_pytest.log
tail -3 $O/j_pytest.log
timeout 300 python tools/v3_peel_profile.py > $O/j_v3_peel_profile.txt 2>&1
tail -2 $O/j_v3_peel_profile.txt
: > $O/j_ab.jsonl
for tm in 4 8 16; do
  SFFTB_V3_TEAM=$tm timeout 300 python tools/gather_ab.py C3 >> $O/j_ab.jsonl 2>> $O/j_err.txt
done
cut -c1-500 $O/j_ab.jsonl
