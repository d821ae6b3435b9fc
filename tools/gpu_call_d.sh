#!/bin/bash
# round 2, call D (1 GPU): parity suite (cluster select, simplified gather), gather sweep (unroll x occupancy),
# pipe-overlap microbenchmark, stage timings of all workloads
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/d_pytest.log 2>&1; echo "pytest rc $?" >> $O/d_pytest.log
tail -3 $O/d_pytest.log
tools/microbench/pipe_overlap > $O/d_pipe_overlap.jsonl 2>&1
cat $O/d_pipe_overlap.jsonl
: > $O/d_sweep.jsonl
for wl in C2 C4 C5 C1; do
  timeout 600 python tools/gather_sweep.py $wl 6 >> $O/d_sweep.jsonl 2>> $O/d_err.txt
done
: > $O/d_ab.jsonl
for wl in C1 C2 C3 C4 C5; do
  timeout 300 python tools/gather_ab.py $wl >> $O/d_ab.jsonl 2>> $O/d_err.txt
done
SFFTB_NO_SELECT_CLUSTER=1 timeout 300 python tools/gather_ab.py C4 >> $O/d_ab.jsonl 2>> $O/d_err.txt
tail -5 $O/d_err.txt
