#!/bin/bash
# round 2, call O (1 GPU): parity suite, bench lines C3/C4/C5/C1 with the child-process CPU leg, default bench
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/o_pytest.log 2>&1; echo "pytest rc $?" >> $O/o_pytest.log
tail -3 $O/o_pytest.log
for wl in C3 C4 C5 C1; do
  timeout 900 python bench.py --workload $wl --no-extras > $O/o_bench_$wl.json 2> $O/o_bench_$wl.err
  cut -c1-200 $O/o_bench_$wl.json
done
timeout 900 python bench.py > $O/o_bench_default.json 2> $O/o_bench_default.err
cut -c1-200 $O/o_bench_default.json
