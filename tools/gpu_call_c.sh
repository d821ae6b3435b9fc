#!/bin/bash
# round 2, call C (1 GPU): gather launch-form sweep; v2 estimation with 256-hit tiles, two CTAs per SM, start-up skew
set -u
mkdir -p gpurun_out
O=gpurun_out
: > $O/c_sweep.jsonl
for wl in C2 C4 C5 C1; do
  timeout 600 python tools/gather_sweep.py $wl >> $O/c_sweep.jsonl 2>> $O/c_err.txt
done
: > $O/c_v2.jsonl
timeout 300 python tools/gather_ab.py C2 20 >> $O/c_v2.jsonl 2>> $O/c_err.txt
for sk in 0 1000 2000 3000 4500; do
  SFFTB_LIB=$PWD/sfft_b200/libsfft_t8.so SFFTB_V2_SKEW=$sk timeout 300 python tools/gather_ab.py C2 20 >> $O/c_v2.jsonl 2>> $O/c_err.txt
done
SFFTB_LIB=$PWD/sfft_b200/libsfft_t8.so SFFTB_V2_SKEW=2000 timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_edge.py tests/test_gpu_v12.py -x -q -k "v2 or C2 or fused or golden or stage" > $O/c_pytest_t8.log 2>&1
tail -3 $O/c_pytest_t8.log
cat $O/c_v2.jsonl | cut -c1-400
