#!/bin/bash
# round 2, call R (1 GPU): parity suite after the side-stream fork of the estimation-row FFT; C1 bench line
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r_pytest.log 2>&1; echo "pytest rc $?" >> $O/r_pytest.log
tail -3 $O/r_pytest.log
timeout 600 python bench.py --workload C1 --no-extras > $O/r_bench_C1.json 2> $O/r_bench_C1.err
cut -c1-250 $O/r_bench_C1.json
