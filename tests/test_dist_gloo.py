"""Host-side logic of the multi-GPU paths on CPU: world_size 2 over gloo.

What is checked without a GPU: the block partitions cover every signal / loop / hit
exactly once; a draw made on rank 0 reaches rank 1 byte for byte; summing per-rank
bucket-spectra buffers whose rows are disjoint reproduces the full array exactly
(this is the one collective of the loop-sharded transform)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import ROOT


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    from sfft_b200 import _lib
    from sfft_b200 import dist as sd

    # 1. the draw of rank 0 becomes everyone's draw
    d = _lib.Draw()
    if rank == 0:
        d.loops = 20
        for i in range(20):
            d.a[i] = 2 * i + 1
            d.ai[i] = 1000 + i
        d.comb_offset[0] = 77
        d.v3_b = 5
    got = sd.broadcast_draw(d, 0)
    assert got.loops == 20 and got.a[7] == 15 and got.ai[19] == 1019 and got.comb_offset[0] == 77 and got.v3_b == 5

    # 1b. the peer-exchange handles (plain bytes) reach every rank in rank order
    blobs = sd.all_gather_bytes(bytes([rank + 1]) * 128)
    assert [b[0] for b in blobs] == list(range(1, world + 1)) and all(len(b) == 128 for b in blobs)
    assert C.sizeof(_lib.PeerHandle) == 128

    # 2. loop-sharded bucket spectra assemble exactly: each rank fills only its own loops
    n, k = 16384, 50
    x, _ = oracle.generate_input(n, k, 3)
    op = oracle.Plan(n, k, 1)
    oracle.seed(17, 9)
    op.exec(x)
    full = op.arr("x_samp").copy()
    loops = op.loops
    lb, le = sd.partition(loops, rank, world)
    part = np.zeros_like(full)
    offs = [min(j, op.loops_loc) * op.B_loc + max(0, j - op.loops_loc) * op.B_est for j in range(loops + 1)]
    part[offs[lb]:offs[le]] = full[offs[lb]:offs[le]]
    t = torch.from_numpy(part.view(np.float64).copy())
    sd.assemble_spectra(t)
    assert t.numpy().tobytes() == full.view(np.float64).tobytes()

    # 3. partitions of signals / hits tile the range
    for total in (0, 1, 7, 20, 4096, 16384000):
        spans = [sd.partition(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))

    # 4. exec_many bookkeeping: counts of the global batch are identical on every rank
    mine = torch.zeros(6, dtype=torch.int64)
    b, e = sd.partition(6, rank, world)
    mine[b:e] = torch.arange(b, e) + 100
    dist.all_reduce(mine)
    assert mine.tolist() == [100, 101, 102, 103, 104, 105]
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
