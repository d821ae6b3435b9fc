"""Multi-GPU paths on real devices (needs >= 2 GPUs; run with `gpurun --gpus 2`):
the loop-sharded transform must reproduce the single-GPU result bit for bit, and the
signal-partitioned batch must reproduce the per-signal results of one GPU."""
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
    for version, n, k in ((1, 1 << 18, 100), (2, 1 << 17, 60), (1, 1 << 22, 50)):
        x, _ = oracle.generate_input(n, k, 77)
        xd = torch.from_numpy(x).cuda()
        plan = m.sfft(n, k, version, strict_parameters=False)
        plan.set_stream(stream.cuda_stream)
        # reference result: the whole transform on this GPU alone, with a broadcast draw
        oracle.seed(17, 5)
        draw = sd.broadcast_draw(plan.draw() if rank == 0 else plan.draw(), 0)
        cnt1 = plan.execute_device(xd, draw)
        loc1, val1 = plan.result()
        o = np.argsort(loc1, kind="stable")
        loc1, val1 = loc1[o], val1[o]
        # sharded over `world` GPUs
        st = sd.ShardedTransform(plan)
        lb, le = st.owned_loops()
        assert (lb, le) == sd.partition(plan.info()["loops_loc"] + plan.info()["loops_est"], rank, world)
        cnt = st.execute(xd, draw)
        loc, val = plan.result()
        if version == 1:
            o = np.argsort(loc, kind="stable")
            assert cnt == cnt1 and np.array_equal(loc[o], loc1)
            assert val[o].tobytes() == val1.tobytes(), "sharded values must be bit-identical"
        else:
            # v2: this rank holds its slice of the pre-filled list, in list order
            b, e = sd.partition(cnt1, rank, world)
            assert cnt == e - b
            full_loc, full_val = np.empty(cnt1, np.int32), np.empty(cnt1, np.complex128)
            # rebuild the unsorted single-GPU list order to compare slices
            cntx = plan.execute_device(xd, draw)
            ul, uv = plan.result()
            cnt = st.execute(xd, draw)
            loc, val = plan.result()
            assert np.array_equal(loc, ul[b:e]) and val.tobytes() == uv[b:e].tobytes()
        plan.close()

    # signal-partitioned batch
    n, k, total = 1 << 16, 50, 6
    plan = m.sfft(n, k, 1, strict_parameters=False)
    plan.set_stream(stream.cuda_stream)
    sigs = [oracle.generate_input(n, k, 200 + i)[0] for i in range(total)]
    oracle.seed(17, 3)
    draws = [sd.broadcast_draw(plan.draw(), 0) for _ in range(total)]
    b, e = sd.partition(total, rank, world)
    local = torch.from_numpy(np.stack(sigs[b:e])).cuda()
    counts = sd.exec_many_sharded(plan, local, draws)
    for i in range(total):
        c = plan.execute_device(torch.from_numpy(sigs[i]).cuda(), draws[i])
        assert c == counts[i], (i, c, counts[i])
    plan.close()
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_two_gpus_nccl(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    port = 29700 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
