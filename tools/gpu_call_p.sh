#!/bin/bash
# round 2, call P (1 GPU): v3 parity + profile after the warp-per-bucket peel; C3 bench line
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests/test_gpu_v3.py tests/test_gpu_fullsize.py tests/test_gpu_edge.py -x -q > $O/p_pytest.log 2>&1; echo "pytest rc $?" >> $O/p_pytest.log
tail -3 $O/p_pytest.log
timeout 300 python tools/v3_peel_profile.py > $O/p_v3_peel_profile.txt 2>&1
tail -2 $O/p_v3_peel_profile.txt
timeout 600 python tools/v3_diag.py 8 > $O/p_v3_diag.txt 2>&1
tail -8 $O/p_v3_diag.txt
timeout 600 python bench.py --workload C3 --no-extras --no-cpu-baseline > $O/p_bench_C3.json 2> $O/p_bench_C3.err
cut -c1-250 $O/p_bench_C3.json
