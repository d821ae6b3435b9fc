#!/bin/bash
# round 2, call S (1 GPU): v1 parity after the vote grid change; C5 / C4 stage timings
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_v12.py tests/test_gpu_fullsize.py tests/test_gpu_edge.py -x -q > $O/s_pytest.log 2>&1; echo "pytest rc $?" >> $O/s_pytest.log
tail -3 $O/s_pytest.log
: > $O/s_ab.jsonl
for wl in C5 C4 C1; do timeout 300 python tools/gather_ab.py $wl >> $O/s_ab.jsonl 2>> $O/s_err.txt; done
cut -c1-600 $O/s_ab.jsonl
