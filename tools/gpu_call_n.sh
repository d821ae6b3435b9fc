#!/bin/bash
# round 2, call N (N GPUs): v1/v2 parity + sharded parity tests, then the bench at N ranks with extras
set -u
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/n${N}_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > $O/n${N}_pytest.log 2>&1; echo "pytest rc $?" >> $O/n${N}_pytest.log
tail -3 $O/n${N}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 3 > $O/n${N}_bench.json 2> $O/n${N}_bench.err
tail -c 1500 $O/n${N}_bench.json
grep -v "^$" $O/n${N}_bench.err | tail -5
