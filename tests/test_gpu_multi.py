"""Multi-GPU paths on real devices (needs >= 2 GPUs; run with `gpurun --gpus 2`):
the loop-sharded transform -- through the NVLink peer exchange and through the NCCL
fallback -- must reproduce the single-GPU result bit for bit, and the signal-partitioned
batch must reproduce the per-signal results of one GPU.  The same check runs inside
`bench.py --gpus N` ("sharded_parity"), which is what the driver sees."""
import os
import sys

import numpy as np
import pytest

from util import ROOT

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import sfft_b200.sfft as m
    from oracle import oracle
    from sfft_b200 import dist as sd

    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    for exchange in ("peer", "nccl"):
        for version, n, k in ((1, 1 << 18, 100), (2, 1 << 17, 60), (2, 1 << 20, 100), (1, 1 << 22, 50)):
            x, _ = oracle.generate_input(n, k, 77)
            xd = torch.from_numpy(x).cuda()
            plan = m.sfft(n, k, version, strict_parameters=False)
            st = sd.ShardedTransform(plan, exchange=exchange)
            if exchange == "peer":
                assert st.exchange == "peer", st.peer_error
            lb, le = st.owned_loops()
            assert (lb, le) == sd.partition(plan.info()["loops_loc"] + plan.info()["loops_est"], rank, world)
            # every rank draws for itself from identically seeded libc state: no broadcast
            st.seed(17, 5)
            draws = [plan.draw() for _ in range(4)]
            for d in draws:          # several transforms: plain first, then graph replays
                ok, compared = sd.sharded_matches_single(plan, st, xd, d)
                assert ok and compared > 0
            # draw=None path: libc state advances identically on every rank
            st.seed(17, 5)
            cnt = st.execute(xd, None)
            loc, val = plan.result()
            off, n_slice = st.slice()
            st.seed(17, 5)
            dist.barrier()
            cnt1 = plan.execute_device(xd, None)
            loc1, val1 = plan.result()
            dist.barrier()
            if version == 1:
                o, o1 = np.argsort(loc, kind="stable"), np.argsort(loc1, kind="stable")
                assert cnt == cnt1 and np.array_equal(loc[o], loc1[o1]) and val[o].tobytes() == val1[o1].tobytes()
            else:
                assert cnt == n_slice and np.array_equal(loc, loc1[off:off + n_slice])
                assert val.tobytes() == val1[off:off + n_slice].tobytes()
            assert st.status()[1] == 0, "a flag wait timed out"
            st.close()
            plan.close()

    # signal-partitioned batch
    n, k, total = 1 << 16, 50, 6
    plan = m.sfft(n, k, 1, strict_parameters=False)
    plan.set_stream(stream.cuda_stream)
    sigs = [oracle.generate_input(n, k, 200 + i)[0] for i in range(total)]
    oracle.seed(17, 3)
    draws = [plan.draw() for _ in range(total)]
    b, e = sd.partition(total, rank, world)
    local = torch.from_numpy(np.stack(sigs[b:e])).cuda()
    counts = sd.exec_many_sharded(plan, local, draws)
    for i in range(total):
        c = plan.execute_device(torch.from_numpy(sigs[i]).cuda(), draws[i])
        assert c == counts[i], (i, c, counts[i])
    plan.close()
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    port = 29700 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
