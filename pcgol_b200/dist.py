"""Multi-GPU sharding of the hot path (one process per GPU, torch.distributed for plumbing).

The path shards only where it splits naturally:
  * batched Nearest/Range: queries split contiguously across ranks, index replicated on every
    GPU (each rank builds it from the same cloud); results land in disjoint slices — no
    data-path collective.
  * scan-pair ICP farm: pairs dealt round-robin to ranks — no collective.
  * one large ICP: the target is split across ranks, the base index is replicated; per
    iteration every rank reduces its slice to 16 float64 sums on the device
    (pcg_icp_partial_dev), the sums are all-reduced (NCCL over NVLink; gloo in the CPU tests)
    and every rank applies the identical tail of Evaluate + Update (pcg_icp_finish), so all
    ranks hold the same transform without a broadcast.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Tuple

import numpy as np

from . import _lib


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced split of range(n): the first n % world ranks get one extra item."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def round_robin(count: int, rank: int, world: int) -> List[int]:
    """Indices of the independent units (scan pairs, clouds) this rank owns."""
    return list(range(rank, count, world))


def chunk_ranges(hist, world: int) -> List[Tuple[int, int]]:
    """Contiguous chunk-id ranges [lo, hi) per rank, balanced by points: rank r starts at the first chunk whose
    cumulative point count reaches r/world of the cloud.  Covers [0, len(hist)) without gaps; ranges may be empty."""
    hist = np.asarray(hist, np.int64)
    cum = np.concatenate([[0], np.cumsum(hist)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world):
        target = (total * r) // world
        c = int(np.searchsorted(cum, target, side="left"))
        cuts.append(min(max(c, cuts[-1]), len(hist)))
    cuts.append(len(hist))
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def sharded_voxelgrid(d_ptr: int, n: int, leaf, chunk, rank: int, world: int, d_out_ptr: int, device: int = 0,
                      stream: int = 0, stride: int = 12, off=(0, 4, 8), group=None, sample_step: Optional[int] = None):
    """One large voxelGrid.Filter over `world` GPUs: the cloud is replicated, rank r filters a range of
    chunk ids (chunks are independent in the reference, voxelgrid.go:102-116) chosen from the chunk histogram so that
    every rank gets about n/world points, and the outputs concatenated in rank order are the reference's output.
    Writes this rank's records to d_out_ptr (capacity n*stride bytes is always enough) and returns
    (n_out_local, counts_of_all_ranks, (cid_lo, cid_hi)); the only collective is the all-gather of the counts."""
    lf = (C.c_float * 3)(*[float(x) for x in leaf])
    ck = (C.c_int64 * 3)(*[int(x) for x in chunk])
    offs = (C.c_int64 * 3)(*off)
    # balancing needs proportions, not counts: large clouds are sampled (every rank computes the same histogram)
    step = int(sample_step) if sample_step else (16 if n >= (1 << 22) else 1)
    n_chunks = C.c_int64(0)
    hist = np.zeros(1 << 16, np.int64)  # enough for most maps; a larger chunk table asks again with its exact size
    _lib.check(_lib.lib.pcg_voxelgrid_chunk_histogram_dev(d_ptr, n, stride, offs, lf, ck, device, step, hist.ctypes.data,
                                                         len(hist), C.byref(n_chunks), stream))
    if n_chunks.value > len(hist):
        hist = np.zeros(n_chunks.value, np.int64)
        _lib.check(_lib.lib.pcg_voxelgrid_chunk_histogram_dev(d_ptr, n, stride, offs, lf, ck, device, step, hist.ctypes.data,
                                                             len(hist), C.byref(n_chunks), stream))
    lo, hi = chunk_ranges(hist[: n_chunks.value], world)[rank]
    n_out = C.c_int64(0)
    _lib.check(_lib.lib.pcg_voxelgrid_filter_chunks_dev(d_ptr, n, stride, offs, lf, ck, lo, hi, device, d_out_ptr,
                                                       C.byref(n_out), stream))
    counts = [n_out.value]
    try:
        import torch
        import torch.distributed as dist
        if world > 1 and dist.is_available() and dist.is_initialized():
            t = torch.tensor([n_out.value], dtype=torch.int64, device="cuda" if dist.get_backend(group) == "nccl" else "cpu")
            gathered = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(gathered, t, group=group)
            counts = [int(g.item()) for g in gathered]
    except ImportError:
        pass
    return n_out.value, counts, (lo, hi)


# ---- one large VoxelGrid with the POINTS sharded (all-to-all by chunk owner) -------------------------
def ordered_bits(x: np.ndarray) -> np.ndarray:
    """Order-preserving uint32 image of float32 values with -0 == +0 (the comparison Go's MinMaxVec3 makes)."""
    x = np.where(x == 0, np.float32(0), x.astype(np.float32))
    b = x.view(np.uint32)
    return np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.uint32)


def decode_minmax_words(words) -> np.ndarray:
    """The six floats {min xyz, max xyz} from the all-reduced (signed MIN) words of pcg_minmax_packed_dev."""
    w = np.asarray(words, np.int64).view(np.uint64) ^ np.uint64(1 << 63)
    w[3:] = ~w[3:]
    out = np.empty(6, np.uint32)
    for k in range(6):
        ob, low = int(w[k] >> np.uint64(32)), int(w[k] & np.uint64(0xffffffff))
        if (k < 3 and int(w[k]) == 0) or (k >= 3 and int(w[k]) == 0xffffffffffffffff):
            out[k] = 0x7fc00000  # a NaN at global point 0 stays (minmax.go:13); which NaN does not matter downstream
            continue
        bits = (ob & 0x7fffffff) if ob & 0x80000000 else (~ob & 0xffffffff)
        if bits == 0 and (low & 1):
            bits = 0x80000000  # the winning zero was a -0
        out[k] = bits
    return out.view(np.float32).copy()


class GpuVgShard:
    """This rank's slice of the cloud on its GPU (a torch uint8 tensor of n * stride bytes) + the C-ABI steps."""

    def __init__(self, records, n: int, stride: int, off, leaf, chunk, device: int, stream: int = 0):
        self.rec, self.n, self.stride, self.off = records, n, stride, tuple(off)
        self.leaf = (C.c_float * 3)(*[float(x) for x in leaf])
        self.chunk = (C.c_int64 * 3)(*[int(x) for x in chunk])
        self.offs = (C.c_int64 * 3)(*off)
        self.device, self.stream = device, stream

    def minmax_packed(self, index_base: int):
        import torch
        acc = torch.empty(6, dtype=torch.int64, device=self.rec.device)
        _lib.check(_lib.lib.pcg_minmax_packed_dev(self.rec.data_ptr(), self.n, self.stride, self.offs, self.device,
                                                 index_base, acc.data_ptr(), self.stream))
        return acc

    def histogram(self, mm6: np.ndarray, sample_step: int) -> np.ndarray:
        n_chunks = C.c_int64(0)
        hist = np.zeros(1 << 16, np.int64)
        for _ in range(2):
            _lib.check(_lib.lib.pcg_voxelgrid_chunk_histogram_mm_dev(
                self.rec.data_ptr(), self.n, self.stride, self.offs, self.leaf, self.chunk, mm6.ctypes.data, self.device,
                sample_step, hist.ctypes.data, len(hist), C.byref(n_chunks), self.stream))
            if n_chunks.value <= len(hist):
                break
            hist = np.zeros(n_chunks.value, np.int64)
        return hist[: n_chunks.value]

    def owner_order(self, mm6: np.ndarray, cuts: np.ndarray):
        import torch
        world = len(cuts) - 1
        perm = torch.empty(max(self.n, 1), dtype=torch.int32, device=self.rec.device)
        self._send = torch.empty(max(self.n, 1) * self.stride, dtype=torch.uint8, device=self.rec.device)
        counts = np.zeros(world, np.int64)
        _lib.check(_lib.lib.pcg_voxelgrid_owner_order_dev(
            self.rec.data_ptr(), self.n, self.stride, self.offs, self.leaf, self.chunk, mm6.ctypes.data,
            cuts.ctypes.data, world, self.device, perm.data_ptr(), counts.ctypes.data, self._send.data_ptr(),
            self.stream))
        return perm[: self.n], counts

    def gather(self, perm):
        return self._send[: self.n * self.stride]  # written by owner_order (the library gathers whole records)

    def filter(self, recv, n_recv: int, mm6: np.ndarray, lo: int, hi: int, out):
        n_out = C.c_int64(0)
        _lib.check(_lib.lib.pcg_voxelgrid_filter_chunks_mm_dev(
            recv.data_ptr(), n_recv, self.stride, self.offs, self.leaf, self.chunk, mm6.ctypes.data, lo, hi, self.device,
            out.data_ptr() if out is not None else None, C.byref(n_out), self.stream))
        return n_out.value


def sharded_voxelgrid_points(shard, index_base: int, rank: int, world: int, out=None, group=None,
                             sample_step: Optional[int] = None, n_total: Optional[int] = None, timings=None):
    """voxelGrid.Filter of ONE cloud whose points are split over the ranks (rank r holds the contiguous slice that
    starts at global index `index_base`; nothing is replicated).  `shard` supplies the local steps (GpuVgShard on a
    GPU; the CPU tests pass a numpy stand-in).  Collectives: one MIN all-reduce of the six min/max words,
    a SUM all-reduce of the chunk histogram, one all-to-all of the records by chunk owner,
    all-gather of the output counts.  Returns (n_out_local, counts_of_all_ranks, (cid_lo, cid_hi), recv_records, out)
    where `out` (allocated here when None is passed) holds this rank's n_out_local output records;
    the ranks' outputs concatenated in rank order are the reference's output (voxelgrid.go:102-133)."""
    import torch
    import torch.distributed as dist

    collective = world > 1 and dist.is_available() and dist.is_initialized()

    def lap(name):  # tuning aid: wall time per step when a dict is passed (adds a device sync per step)
        if timings is not None:
            import time
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            now = time.perf_counter()
            timings[name] = timings.get(name, 0.0) + 1e3 * (now - timings.get("_t", now))
            timings["_t"] = now

    lap("start")
    # 1. MinMaxVec3 of the whole cloud: first occurrence of the extreme value wins across ranks (one all-reduce)
    acc = shard.minmax_packed(index_base)
    if collective:
        dist.all_reduce(acc, op=dist.ReduceOp.MIN, group=group)
    mm6 = decode_minmax_words(acc.cpu().numpy())
    lap("minmax")
    # 2. chunk ranges balanced by points
    total = n_total if n_total is not None else shard.n * world
    step = int(sample_step) if sample_step else (16 if total >= (1 << 22) else 1)
    hist = shard.histogram(mm6, step)
    if collective:
        t = torch.from_numpy(hist.copy()).to(acc.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        hist = t.cpu().numpy()
    ranges = chunk_ranges(hist, world)
    cuts = np.array([r[0] for r in ranges] + [len(hist)], np.int64)
    lo, hi = ranges[rank]
    lap("histogram")
    # 3. records to their chunk's owner (whole records: the filter keeps the first member's fields)
    perm, counts = shard.owner_order(mm6, cuts)
    send = shard.gather(perm)
    lap("owner_order")
    if collective:
        cin = torch.from_numpy(counts.copy()).to(acc.device)
        cout = torch.empty_like(cin)
        dist.all_to_all_single(cout, cin, group=group)
        recv_counts = cout.cpu().numpy()
        recv = torch.empty(int(recv_counts.sum()) * shard.stride, dtype=send.dtype, device=send.device)
        dist.all_to_all_single(recv, send, output_split_sizes=[int(c) * shard.stride for c in recv_counts],
                               input_split_sizes=[int(c) * shard.stride for c in counts], group=group)
    else:
        recv_counts = counts
        recv = send
    n_recv = int(recv_counts.sum())
    lap("all_to_all")
    # 4. the owner filters what it received: sources arrive in rank order = global point order
    if out is None:  # a rank's chunks may hold more points than its slice: sized after the exchange
        out = torch.empty(max(1, n_recv) * shard.stride, dtype=torch.uint8, device=recv.device)
    n_out = shard.filter(recv, n_recv, mm6, lo, hi, out)
    lap("filter")
    all_counts = [n_out]
    if collective:
        t = torch.tensor([n_out], dtype=torch.int64, device=acc.device)
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t, group=group)
        all_counts = [int(c) for c in torch.cat(gathered).cpu().tolist()]  # one device -> host copy, not one per rank
    lap("counts")
    return n_out, all_counts, (lo, hi), recv, out


def sharded_icp_fit(partial_fn: Callable[[np.ndarray, bool], "object"], params: _lib.IcpParams,
                    group=None, all_reduce: Optional[Callable] = None):
    """PointToPointICPGradient.Fit (icp.go:23-67) over a target split across ranks.

    partial_fn(trans16, first) returns this rank's 16 float64 sums for the current transform as a
    torch tensor (on the GPU for NCCL, on the CPU for gloo): {Value, SumW, G0..G5, R, nPairs, 0...}.
    Returns (status, trans16, Evaluated, num_iteration); identical on every rank.
    """
    import torch
    import torch.distributed as dist

    trans = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1], np.float32)  # icp.go:47
    it = C.c_int32(0)
    conv = C.c_int32(0)
    ev = _lib.Evaluated()
    num_iteration = 0
    status = _lib.OK
    while True:
        sums = partial_fn(trans, num_iteration == 0)
        if all_reduce is not None:
            sums = all_reduce(sums)
        elif dist.is_available() and dist.is_initialized():
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        host = sums.detach().to("cpu", torch.float64).numpy() if hasattr(sums, "detach") else np.asarray(sums)
        host = np.ascontiguousarray(host, np.float64)
        num_iteration += 1  # icp.go:50
        status = _lib.lib.pcg_icp_finish(host.ctypes.data, C.byref(params), C.byref(it), trans.ctypes.data,
                                         C.byref(ev), C.byref(conv))
        if status != _lib.OK or conv.value:
            break
    return status, trans, ev, num_iteration


def make_gpu_partial(index, d_target_ptr: int, n: int, max_dist: float, stream: int = 0, stride: int = 12,
                     off=(0, 4, 8)):
    """partial_fn for sharded_icp_fit backed by pcg_icp_partial_dev (the rank's target slice is device resident)."""
    import torch

    out = torch.zeros(16, dtype=torch.float64, device="cuda")
    offs = (C.c_int64 * 3)(*off)
    order = None
    if n >= (1 << 14) and len(index) > 0:  # Morton visit order, computed once: the target only moves rigidly
        order = torch.empty(n, dtype=torch.int32, device="cuda")
        _lib.check(_lib.lib.pcg_query_order_dev(index._h, d_target_ptr, n, stride, offs, order.data_ptr(), stream))

    def partial(trans: np.ndarray, first: bool):
        _lib.check(_lib.lib.pcg_icp_partial_dev(index._h, d_target_ptr, n, stride, offs, max_dist, trans.ctypes.data,
                                                1 if first else 0, order.data_ptr() if order is not None else None,
                                                out.data_ptr(), stream))
        return out

    return partial


def sharded_icp_fit_device(index, d_target_ptr: int, n: int, params: _lib.IcpParams, group=None, stream: int = 0,
                           stride: int = 12, off=(0, 4, 8), poll_every: int = 5):
    """The same Fit with the loop resident on the device (pcg_icp_shard_*): per iteration the rank's slice is
    reduced to 16 float64 on the GPU, all-reduced in stream order (NCCL over NVLink) and the tail of Evaluate +
    Update runs on the device - no host round trip inside the loop.  `stream` must be torch's current stream (the
    collective is ordered against it).  The end of the loop is polled every `poll_every` iterations (0 = never:
    all MaxIteration iterations are enqueued; finished ones fall through).
    Returns (status, trans16, IcpStat); identical on every rank."""
    import torch
    import torch.distributed as dist

    offs = (C.c_int64 * 3)(*off)
    sh = C.c_void_p()
    _lib.check(_lib.lib.pcg_icp_shard_new(index._h, d_target_ptr, n, stride, offs, C.byref(params), stream, C.byref(sh)))
    try:
        buf = torch.zeros(16, dtype=torch.float64, device="cuda")
        max_iter = params.max_iteration or 20  # updater.go:31-33
        trans = np.zeros(16, np.float32)
        stat = _lib.IcpStat()
        done = C.c_int32(0)
        status = _lib.OK
        collective = dist.is_available() and dist.is_initialized()
        for it in range(max_iter):
            _lib.check(_lib.lib.pcg_icp_shard_partial(sh, buf.data_ptr(), stream))
            if collective:
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
            _lib.check(_lib.lib.pcg_icp_shard_finish(sh, buf.data_ptr(), stream))
            if poll_every and (it + 1) % poll_every == 0 and it + 1 < max_iter:
                status = _lib.lib.pcg_icp_shard_result(sh, trans.ctypes.data, C.byref(stat), C.byref(done), stream)
                if done.value:
                    return status, trans, stat
        status = _lib.lib.pcg_icp_shard_result(sh, trans.ctypes.data, C.byref(stat), C.byref(done), stream)
        return status, trans, stat
    finally:
        _lib.lib.pcg_icp_shard_free(sh)
