#!/bin/bash
# pipelined v2 estimation kernel (one component per thread): GPU suite, then C2 with and without it
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/v_pytest.log 2>&1; echo "pytest rc $?" >> $O/v_pytest.log
tail -3 $O/v_pytest.log
timeout 300 python bench.py --no-extras --no-cpu-baseline > $O/v_bench_C2_pipe.json 2> $O/v_bench.err
cut -c1-200 $O/v_bench_C2_pipe.json
SFFTB_NO_V2_PIPE=1 timeout 300 python bench.py --no-extras --no-cpu-baseline > $O/v_bench_C2_nopipe.json 2>> $O/v_bench.err
cut -c1-200 $O/v_bench_C2_nopipe.json
python - <<'PY'
import json
for t in ("pipe", "nopipe"):
    try:
        d = json.load(open(f"gpurun_out/v_bench_C2_{t}.json")); print(t, d["ms_per_step"], d["stage_ms"])
    except Exception as e: print(t, e)
PY
tail -3 $O/v_bench.err
