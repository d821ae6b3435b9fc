#!/bin/bash
# round 2, call I (1 GPU): v3 + full-size parity on the current tree, v3 profile
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_v3.py tests/test_gpu_fullsize.py tests/test_gpu_v12.py -x -q > $O/i_pytest.log 2>&1; echo "pytest rc $?" >> $O/i_pytest.log
tail -3 $O/i_pytest.log
timeout 300 python tools/v3_peel_profile.py > $O/i_v3_peel_profile.txt 2>&1
tail -2 $O/i_v3_peel_profile.txt
timeout 300 python tools/gather_ab.py C3 > $O/i_ab.jsonl 2>> $O/i_err.txt
cut -c1-500 $O/i_ab.jsonl
