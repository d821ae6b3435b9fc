"""Multi-GPU plumbing (one process per GPU, torch.distributed).

Two ways the transform path shards (SURVEY 8e / DESIGN.md "Multi-GPU"):

* independent signals (sfft_exec_many): block-partition the signals over ranks, no
  data-path collective -- `partition`, `exec_many_sharded`;
* one large v1/v2 signal: every rank bucketises its own block of loops, ONE
  all-reduce (a sum in which each element has exactly one non-zero contributor)
  completes the bucket spectra everywhere, selection/voting are replicated and the
  estimation of v2's pre-filled list is sliced -- `ShardedTransform`.

v3 has no loop structure to shard: replicas only.

The permutations must be identical on every rank: rank 0 draws them from libc
random()/drand48() in the reference's order and broadcasts the draw.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def partition(total, rank, world):
    """Block partition used for signals, loops and hit slices: [begin, end)."""
    return total * rank // world, total * (rank + 1) // world


class _DevArray:
    """Zero-copy view of device memory for torch (via __cuda_array_interface__)."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {
            "shape": (int(count),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_view(ptr, count, dtype, device):
    typestr = {torch.float64: "<f8", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_DevArray(ptr, count, typestr), device=device)


def _comm_device(group=None):
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" \
        else torch.device("cpu")


def broadcast_draw(draw, src=0, group=None):
    """Make rank `src`'s sfftb_draw the draw of every rank."""
    raw = np.frombuffer(bytes(draw), dtype=np.uint8).copy()
    t = torch.from_numpy(raw).to(_comm_device(group))
    dist.broadcast(t, src=src, group=group)
    out = _lib.Draw.from_buffer_copy(t.cpu().numpy().tobytes())
    return out


def assemble_spectra(partial, group=None):
    """Sum of per-rank bucket-spectra buffers in which every row is non-zero on exactly
    one rank: an exact all-gather with uneven row blocks."""
    dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial


class ShardedTransform:
    """One v1/v2 transform of a device-resident signal spread over the ranks of `group`."""

    def __init__(self, plan, group=None):
        self.plan = plan
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.L = _lib.load()
        if plan.version == 3:
            raise ValueError("sFFT v3 has no independent loops to shard (replicas only)")
        ptr, cnt = C.c_void_p(), C.c_longlong()
        if self.L.sfftb_shard_spectra(plan.sfft_plan, C.byref(ptr), C.byref(cnt)):
            raise RuntimeError(_lib.last_error())
        self.spectra = device_view(ptr.value, cnt.value, torch.float64,
                                   torch.device("cuda", torch.cuda.current_device()))

    def owned_loops(self):
        b, e = C.c_int(), C.c_int()
        if self.L.sfftb_shard_loops(self.plan.sfft_plan, self.rank, self.world, C.byref(b), C.byref(e)):
            raise RuntimeError(_lib.last_error())
        return b.value, e.value

    def execute(self, x, draw=None, sync=True):
        """x: CUDA complex128[n] holding the SAME signal on every rank.  Returns this
        rank's number of results (v1: all hits, replicated; v2: its slice of the list)."""
        if draw is None:
            draw = self.plan.draw() if self.rank == 0 else _lib.Draw()
            draw = broadcast_draw(draw, 0, self.group)
        if self.L.sfftb_shard_bucketize(self.plan.sfft_plan, C.c_void_p(x.data_ptr()), C.byref(draw),
                                        self.rank, self.world):
            raise RuntimeError(_lib.last_error())
        assemble_spectra(self.spectra, self.group)
        res = _lib.Result()
        if self.L.sfftb_shard_finish(self.plan.sfft_plan, self.rank, self.world, C.byref(res),
                                     1 if sync else 0):
            raise RuntimeError(_lib.last_error())
        return int(res.count) if sync else None


def exec_many_sharded(plan, signals, draws, group=None):
    """signals: CUDA complex128[num_local, n], THIS rank's block of a global batch whose
    draws (one per global signal, drawn in global signal order) are `draws`.  Returns
    the per-signal result counts of the whole batch on every rank."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    total = len(draws)
    b, e = partition(total, rank, world)
    assert signals.shape[0] == e - b
    counts = plan.execute_many_device(signals, draws[b:e], sync=True) if e > b else []
    mine = torch.zeros(total, dtype=torch.int64, device=_comm_device(group))
    if e > b:
        mine[b:e] = torch.tensor(counts, dtype=torch.int64, device=mine.device)
    dist.all_reduce(mine, op=dist.ReduceOp.SUM, group=group)      # bookkeeping only
    return mine.cpu().tolist()
