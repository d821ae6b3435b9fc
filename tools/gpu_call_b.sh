#!/bin/bash
# round 2, call B (2+ GPUs): sharded-transform parity test, then the bench at N ranks with extras
set -u
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/b_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > $O/b_pytest.log 2>&1; echo "pytest rc $?" >> $O/b_pytest.log
tail -5 $O/b_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 > $O/b_bench_N$N.json 2> $O/b_bench_N$N.err
tail -c 3000 $O/b_bench_N$N.json
tail -5 $O/b_bench_N$N.err
