#!/bin/bash
# Comb stage forked onto the side stream: v1/v2 suites, then the C2 / C2b lines
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/t_pytest.log 2>&1; echo "pytest rc $?" >> $O/t_pytest.log
tail -3 $O/t_pytest.log
timeout 600 python bench.py --no-extras --no-cpu-baseline > $O/t_bench_C2.json 2> $O/t_bench.err
cut -c1-200 $O/t_bench_C2.json
timeout 300 python bench.py --workload C2b --no-extras --no-cpu-baseline > $O/t_bench_C2b.json 2>> $O/t_bench.err
cut -c1-200 $O/t_bench_C2b.json
tail -3 $O/t_bench.err
