"""Multi-GPU plumbing (one process per GPU, torch.distributed).

Two ways the transform path shards (SURVEY 8e / DESIGN.md "Multi-GPU"):

* independent signals (sfft_exec_many): block-partition the signals over ranks, no
  data-path collective -- `partition`, `exec_many_sharded`;
* one large v1/v2 signal: every rank bucketises its own block of loops, ONE exchange
  completes the bucket spectra everywhere, selection/voting are replicated and the
  estimation of v2's pre-filled list is sliced -- `ShardedTransform`.  The exchange is
  the library's own NVLink peer exchange (stores into CUDA-IPC-mapped peer buffers +
  flags, inside the transform's CUDA graph; include/sfft.h "NVLink peer exchange");
  `exchange="nccl"` selects the portable fallback, one torch.distributed all-reduce in
  which every element has exactly one non-zero contributor.

v3 has no loop structure to shard: replicas only.

The permutations must be identical on every rank.  `ShardedTransform.seed(s, s48)` seeds
libc identically everywhere, after which every rank draws the same sequence by itself
(no per-transform broadcast, no host synchronisation); `broadcast_draw` is there for
callers that would rather draw on one rank.
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
    typestr = {torch.float64: "<f8", torch.int32: "<i4", torch.int64: "<i8"}[dtype]
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


def all_gather_bytes(blob, group=None):
    """Every rank's `blob` (bytes of equal length), as a list indexed by rank."""
    world = dist.get_world_size(group)
    mine = torch.from_numpy(np.frombuffer(blob, dtype=np.uint8).copy()).to(_comm_device(group))
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [t.cpu().numpy().tobytes() for t in out]


class ShardedTransform:
    """One v1/v2 transform of a device-resident signal spread over the ranks of `group`.

    Stream contract: the plan is bound to `self.stream` (torch's current stream at
    construction, or a fresh one when that is the legacy default stream); the signal
    passed to `execute` must be ready on that stream, and the result is ready on it."""

    def __init__(self, plan, group=None, exchange="peer"):
        if plan.version == 3:
            raise ValueError("sFFT v3 has no independent loops to shard (replicas only)")
        self.plan = plan
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.L = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device())
        cur = torch.cuda.current_stream()
        # handle 0 (the legacy default stream) reads as "the plan's own stream" in the C ABI
        self.stream = cur if cur.cuda_stream != 0 else torch.cuda.Stream(device=self.device)
        plan.set_stream(self.stream.cuda_stream)
        self.exchange = "nccl"
        self.peer_error = None
        if exchange == "peer":
            self._attach_peers()
        elif exchange != "nccl":
            raise ValueError("exchange must be 'peer' or 'nccl'")

    # ---- setup ----
    def _attach_peers(self):
        P = self.plan.sfft_plan
        mine = _lib.PeerHandle()
        ok = self.L.sfftb_shard_export(P, C.byref(mine)) == 0
        err = None if ok else _lib.last_error()
        blobs = all_gather_bytes(bytes(mine), self.group)
        if ok:
            arr = (_lib.PeerHandle * self.world)(*[_lib.PeerHandle.from_buffer_copy(b) for b in blobs])
            ok = self.L.sfftb_shard_attach(P, self.rank, self.world, arr) == 0
            if not ok:
                err = _lib.last_error()
        # every rank must take the same path
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=_comm_device(self.group))
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 1:
            self.exchange = "peer"
        else:
            self.peer_error = err or "a peer could not map the buffers"
            if ok:
                dist.barrier(group=self.group)
                self.L.sfftb_shard_detach(P)
        dist.barrier(group=self.group)

    def seed(self, s=17, s48=12345):
        """srand(s); srand48(s48) on this rank -- call it with the same arguments on every
        rank, then `execute(x)` draws the reference's permutations identically everywhere."""
        libc = C.CDLL(None)
        libc.srand(C.c_uint(s))
        libc.srand48(C.c_long(s48))

    def owned_loops(self):
        b, e = C.c_int(), C.c_int()
        if self.L.sfftb_shard_loops(self.plan.sfft_plan, self.rank, self.world, C.byref(b), C.byref(e)):
            raise RuntimeError(_lib.last_error())
        return b.value, e.value

    # ---- one transform ----
    def execute(self, x, draw=None, sync=True):
        """x: CUDA complex128[n] holding the SAME signal on every rank.  Returns this
        rank's number of results (v1: all hits, replicated; v2: its slice of the list)."""
        P = self.plan.sfft_plan
        res = _lib.Result()
        if self.exchange == "peer":
            if self.L.sfftb_shard_exec(P, C.c_void_p(x.data_ptr()), C.byref(draw) if draw is not None else None,
                                       C.byref(res), 1 if sync else 0):
                raise RuntimeError(_lib.last_error())
            self.plan._last = res
            return int(res.count) if sync else None
        if draw is None:
            draw = self.plan.draw()        # identical on every rank when libc was seeded identically
        if self.L.sfftb_shard_bucketize(P, C.c_void_p(x.data_ptr()), C.byref(draw), self.rank, self.world):
            raise RuntimeError(_lib.last_error())
        # the buffer can move when a batch call grew the plan's scratch: ask every time
        ptr, cnt = C.c_void_p(), C.c_longlong()
        if self.L.sfftb_shard_spectra(P, C.byref(ptr), C.byref(cnt)):
            raise RuntimeError(_lib.last_error())
        spectra = device_view(ptr.value, cnt.value, torch.float64, self.device)
        with torch.cuda.stream(self.stream):      # NCCL orders itself against torch's current stream
            assemble_spectra(spectra, self.group)
        if self.L.sfftb_shard_finish(P, self.rank, self.world, C.byref(res), 1 if sync else 0):
            raise RuntimeError(_lib.last_error())
        self.plan._last = res
        return int(res.count) if sync else None

    def slice(self):
        """(offset, count): the entries of the single-GPU result list this rank produced in the
        last transform (v1: the whole list)."""
        off, cnt = C.c_longlong(), C.c_longlong()
        if self.L.sfftb_shard_slice(self.plan.sfft_plan, self.rank, self.world, C.byref(off), C.byref(cnt)):
            raise RuntimeError(_lib.last_error())
        return off.value, cnt.value

    def status(self):
        """(transforms completed, flag waits that timed out) of the peer exchange."""
        if self.exchange != "peer":
            return 0, 0
        e, t = C.c_longlong(), C.c_longlong()
        if self.L.sfftb_shard_status(self.plan.sfft_plan, C.byref(e), C.byref(t)):
            raise RuntimeError(_lib.last_error())
        return e.value, t.value

    def close(self):
        if self.exchange == "peer":
            self.plan.synchronize()
            dist.barrier(group=self.group)        # nobody stores into a buffer that is being unmapped
            self.L.sfftb_shard_detach(self.plan.sfft_plan)
            dist.barrier(group=self.group)
            self.exchange = "closed"


def sharded_matches_single(plan, st, x, draw):
    """Driver-visible parity check of the loop-sharded transform: run the same (signal, draw)
    sharded and on this GPU alone and compare this rank's part of the result bit for bit, on
    the device.  Returns (ok on every rank, number of entries compared on this rank)."""
    cnt = st.execute(x, draw, sync=True)
    loc_s, val_s = plan.result_device()
    loc_s, val_s = loc_s.clone(), val_s.clone()
    off, n_slice = st.slice()
    cnt1 = plan.execute_device(x, draw, sync=True)
    loc_1, val_1 = plan.result_device()
    # a peer that is already in its next sharded transform would store into this rank's
    # spectra buffer while the single-GPU transform above still used it
    dist.barrier(group=st.group)
    ok = cnt == n_slice and off + n_slice <= cnt1
    if ok and plan.version == 1:
        # v1 lists come out in atomic-append order: compare as sorted sets
        ok = cnt == cnt1
        if ok:
            o_s, o_1 = torch.argsort(loc_s), torch.argsort(loc_1)
            bits_s, bits_1 = val_s.view(torch.int64).view(-1, 2), val_1.view(torch.int64).view(-1, 2)
            ok = bool(torch.equal(loc_s[o_s], loc_1[o_1])) and bool(torch.equal(bits_s[o_s], bits_1[o_1]))
    elif ok:
        ok = bool(torch.equal(loc_s, loc_1[off:off + n_slice])) and \
            bool(torch.equal(val_s.view(torch.int64), val_1[off:off + n_slice].view(torch.int64)))
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=_comm_device(st.group))
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=st.group)
    return bool(flag.item()), int(n_slice)


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
