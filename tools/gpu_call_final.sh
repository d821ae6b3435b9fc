#!/bin/bash
# round 2, final validation (1 GPU), as the driver does it: GPU test-suite, smoke, default bench, reference arm
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/z_pytest.log 2>&1; echo "pytest rc $?" >> $O/z_pytest.log
tail -3 $O/z_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/z_smoke.log 2>&1; echo "smoke rc $?" >> $O/z_smoke.log
tail -5 $O/z_smoke.log
( time timeout 900 python bench.py ) > $O/z_bench.json 2> $O/z_bench.err
cut -c1-300 $O/z_bench.json
timeout 600 python bench.py --impl reference > $O/z_bench_reference.json 2>> $O/z_bench.err
cut -c1-300 $O/z_bench_reference.json
tail -4 $O/z_bench.err
