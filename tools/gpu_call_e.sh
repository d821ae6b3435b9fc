#!/bin/bash
# round 2, call E (1 GPU): pipe microbenchmark (warp-split), plan-builder step timing, default bench as the driver runs it
set -u
mkdir -p gpurun_out
O=gpurun_out
tools/microbench/pipe_overlap > $O/e_pipe_overlap.jsonl 2>&1
cat $O/e_pipe_overlap.jsonl
for wl in C2 C3 C4; do
  echo "== $wl" >> $O/e_plan_timing.txt
  SFFTB_PLAN_TIMING=1 timeout 300 python tools/gather_ab.py $wl 3 2>> $O/e_plan_timing.txt | cut -c1-300
done
cat $O/e_plan_timing.txt
( time timeout 900 python bench.py ) > $O/e_bench_default.json 2> $O/e_bench_default.err
tail -c 2500 $O/e_bench_default.json
tail -5 $O/e_bench_default.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/e_bench_reference.json 2>> $O/e_bench_default.err
cat $O/e_bench_reference.json | cut -c1-1500
