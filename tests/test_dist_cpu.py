"""World-size-2 gloo tests of the multi-GPU host logic (no GPU needed): sharding helpers and
the sharded-ICP loop (all-reduce of the 16 sums + pcg_icp_finish on every rank)."""
import os
import socket

import numpy as np
import pytest


def test_shard_bounds_cover_range():
    from pcgol_b200.dist import round_robin, shard_bounds

    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert round_robin(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((round_robin(10, r, 4) for r in range(4)), [])) == list(range(10))


def _partial_numpy(orc, base_search, base, target, max_dist):
    """CPU stand-in for pcg_icp_partial_dev: float64 sums of the nine Evaluate terms over a target slice."""
    import torch

    def partial(trans, first):
        t = target if first else orc.mat4_transform(trans, target)
        ids, dsq = base_search.nearest(t, max_dist)
        m = ids >= 0
        pt = t[m].astype(np.float32)
        pb = base[ids[m]].astype(np.float32)
        s = np.zeros(16, np.float64)
        s[0] = dsq[m].astype(np.float64).sum()
        s[1] = m.sum()
        d = (pt - pb).astype(np.float32)
        s[2:5] = d.astype(np.float64).sum(axis=0)
        x0, y0, z0 = pt[:, 0], pt[:, 1], pt[:, 2]
        x1, y1, z1 = pb[:, 0], pb[:, 1], pb[:, 2]
        f32 = np.float32
        s[5] = (f32(z0 * y1) - f32(y0 * z1)).astype(f32).astype(np.float64).sum()
        s[6] = (f32(x0 * z1) - f32(z0 * x1)).astype(f32).astype(np.float64).sum()
        s[7] = (f32(y0 * x1) - f32(x0 * y1)).astype(f32).astype(np.float64).sum()
        s[8] = ((x0 * x0 + y0 * y0).astype(f32) + z0 * z0).astype(f32).astype(np.float64).sum()
        s[9] = m.sum()
        return torch.from_numpy(s)

    return partial


def _make_problem():
    rng = np.random.default_rng(42)
    base = (rng.random((4000, 3)) * np.array([10, 10, 2])).astype(np.float32)
    ang = 0.03
    R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    target = (base @ R.T + np.array([0.05, -0.04, 0.02])).astype(np.float32)
    return base, target


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from oracle import oracle as orc
    from pcgol_b200 import _lib
    from pcgol_b200.dist import shard_bounds, sharded_icp_fit

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        base, target = _make_problem()
        lo, hi = shard_bounds(len(target), rank, world)
        search = orc.Search(base, "kdtree")
        p = _lib.IcpParams()
        p.max_dist, p.mode = 0.5, _lib.ICP_FAST
        status, trans, ev, iters = sharded_icp_fit(_partial_numpy(orc, search, base, target[lo:hi], 0.5), p)
        q.put((rank, status, trans.tobytes(), iters, float(ev.value)))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.timeout(300)
def test_sharded_icp_world2_gloo(oracle):
    import torch.multiprocessing as mp

    from pcgol_b200 import _lib
    from pcgol_b200.dist import sharded_icp_fit

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=240) for _ in procs)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    (r0, s0, t0, i0, v0), (r1, s1, t1, i1, v1) = res
    assert s0 == s1 == 0 and i0 == i1 and t0 == t1 and v0 == v1  # every rank holds the identical transform

    # single-process run of the same loop (world 1): float64 partial sums differ only by association
    base, target = _make_problem()
    search = oracle.Search(base, "kdtree")
    p = _lib.IcpParams()
    p.max_dist, p.mode = 0.5, _lib.ICP_FAST
    status, trans, ev, iters = sharded_icp_fit(_partial_numpy(oracle, search, base, target, 0.5), p)
    assert status == 0 and iters == i0
    np.testing.assert_allclose(np.frombuffer(t0, np.float32), trans, rtol=0, atol=1e-6)

    # and against the sequential reference restatement with float64 accumulation
    rc, etrans, _, eit = oracle.icp_fit(search, target, oracle.icp_params(0.5, f64_accumulate=True))
    assert rc == oracle.OK and eit == iters
    np.testing.assert_allclose(trans, etrans, rtol=0, atol=1e-6)


def test_chunk_ranges_partition_the_chunk_table():
    # host logic of the sharded VoxelGrid: contiguous, gap-free, balanced by points
    from pcgol_b200 import dist as pdist

    rng = np.random.default_rng(0)
    for trial in range(50):
        nchunks = int(rng.integers(1, 400))
        hist = rng.integers(0, 5000, nchunks) * (rng.random(nchunks) < 0.6)
        for world in (1, 2, 3, 4, 8):
            r = pdist.chunk_ranges(hist, world)
            assert len(r) == world and r[0][0] == 0 and r[-1][1] == nchunks
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert all(lo <= hi for lo, hi in r)
            pts = [int(hist[lo:hi].sum()) for lo, hi in r]
            assert sum(pts) == int(hist.sum())
            if hist.sum() > 0:  # no rank exceeds its share by more than the largest chunk
                assert max(pts) <= hist.sum() / world + hist.max()
    assert pdist.chunk_ranges([0, 0, 0], 2) == [(0, 0), (0, 3)] or pdist.chunk_ranges([0, 0, 0], 2)[-1][1] == 3


# ---- point-sharded VoxelGrid: the plan (bounds, ranges, exchange) on CPU ranks -------------------------------------
class _NumpyVgShard:
    """CPU stand-in for dist.GpuVgShard: the same steps in numpy float32 (Go's arithmetic: every op rounded)."""

    def __init__(self, xyz, leaf, chunk):
        import torch
        self.torch = torch
        self.xyz = np.ascontiguousarray(xyz, np.float32)
        self.n, self.stride = len(xyz), 12
        self.leaf = np.float32(leaf)
        self.chunk = chunk

    def minmax_packed(self, index_base):
        # the packing of pcg_minmax_packed_dev, restated: value bits | global index << 1 | sign of a zero; maxima
        # complemented so that one signed MIN all-reduce serves all six words
        from pcgol_b200.dist import ordered_bits
        out = np.empty(6, np.uint64)
        g = (np.arange(self.n) + index_base).astype(np.uint64)
        for k in range(3):
            col = self.xyz[:, k]
            ok = col == col
            o = ordered_bits(col[ok]).astype(np.uint64) << np.uint64(32)
            nz = (col[ok].view(np.uint32) == np.uint32(0x80000000)).astype(np.uint64)
            out[k] = (o | (g[ok] << np.uint64(1)) | nz).min() if ok.any() else np.uint64(0xffffffffffffffff)
            mx = (o | ((np.uint64(0x7fffffff) - g[ok]) << np.uint64(1)) | nz).max() if ok.any() else np.uint64(0)
            if index_base == 0 and self.n and not ok[0]:
                out[k], mx = np.uint64(0), np.uint64(0xffffffffffffffff)
            out[3 + k] = ~mx
        return self.torch.from_numpy((out ^ np.uint64(1 << 63)).view(np.int64).copy())

    def _cids(self, mm6):
        vmin, vmax = mm6[:3], mm6[3:]
        size = (vmax - vmin).astype(np.float32)
        cs = np.minimum((self.leaf * np.float32(self.chunk)).astype(np.float32), (size + self.leaf).astype(np.float32))
        ncx, ncy, _ = ((size / cs).astype(np.float32).astype(np.int64) + 1)
        c = ((self.xyz - vmin).astype(np.float32) / cs).astype(np.float32).astype(np.int64)
        return (c[:, 2] * ncy + c[:, 1]) * ncx + c[:, 0], int(np.prod((size / cs).astype(np.float32).astype(np.int64) + 1))

    def histogram(self, mm6, sample_step):
        cid, n_chunks = self._cids(mm6)
        return np.bincount(cid, minlength=n_chunks).astype(np.int64)

    def owner_order(self, mm6, cuts):
        cid, _ = self._cids(mm6)
        owner = np.searchsorted(cuts[1:-1], cid, side="right")
        perm = np.argsort(owner, kind="stable")
        return self.torch.from_numpy(perm.astype(np.int32)), np.bincount(owner, minlength=len(cuts) - 1).astype(np.int64)

    def gather(self, perm):
        return self.torch.from_numpy(self.xyz[perm.numpy()].reshape(-1).view(np.uint8).copy())

    def filter(self, recv, n_recv, mm6, lo, hi, out):
        self.received = recv.numpy().view(np.float32).reshape(-1, 3).copy()
        self.mm6 = mm6.copy()
        return n_recv  # the plan is what this test checks; the filter itself is covered on the GPU


def _vg_cloud():
    rng = np.random.default_rng(7)
    xyz = (rng.random((9001, 3)) * np.array([6.0, 5.0, 1.5]) - np.array([1.0, 2.0, 0.25])).astype(np.float32)
    xyz[100, 1] = xyz[:, 1].min()        # the extreme value twice: the first occurrence is what counts
    xyz[7000, 0] = xyz[:, 0].max()       # ... and once in each rank's slice
    xyz[20, 0] = xyz[:, 0].max()
    xyz[:, 2] = np.abs(xyz[:, 2])
    xyz[5000, 2] = np.float32(-0.0)      # the minimum is a zero: -0 first (slice 1), +0 later - the sign must survive
    xyz[8000, 2] = np.float32(0.0)
    return xyz


def _vg_worker(rank, world, port, q):
    import torch.distributed as dist

    from pcgol_b200.dist import shard_bounds, sharded_voxelgrid_points

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        xyz = _vg_cloud()
        lo, hi = shard_bounds(len(xyz), rank, world)
        shard = _NumpyVgShard(xyz[lo:hi], 0.1, 8)
        n_out, counts, (clo, chi), _, _ = sharded_voxelgrid_points(shard, lo, rank, world, n_total=len(xyz))
        q.put((rank, shard.mm6.tobytes(), clo, chi, shard.received.tobytes(), counts))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_point_sharded_voxelgrid_plan_world2_gloo(oracle):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_vg_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=240) for _ in procs)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    xyz = _vg_cloud()
    # the reduced bounds are MinMaxVec3 of the whole cloud, bit for bit (first occurrence across ranks, zero sign)
    mn, mx = oracle.minmax(xyz.view(np.uint8).reshape(-1), 12, (0, 4, 8))
    for r in res:
        assert r[1] == np.concatenate([mn, mx]).astype(np.float32).tobytes()
    # contiguous chunk ranges; every rank received exactly the points of its chunks, in global point order
    assert res[0][2] == 0 and res[0][3] == res[1][2]
    whole = _NumpyVgShard(xyz, 0.1, 8)
    cid, n_chunks = whole._cids(np.frombuffer(res[0][1], np.float32))
    assert res[1][3] == n_chunks
    for r in res:
        mine = xyz[(cid >= r[2]) & (cid < r[3])]
        assert r[4] == mine.tobytes()
    assert res[0][5] == res[1][5] and sum(res[0][5]) == len(xyz)


def test_minmax_words_decode_nan_at_point_zero_and_zero_sign():
    """The reduced words carry everything MinMaxVec3 needs: a NaN at global point 0 is kept (minmax.go:13), other NaNs
    never win, and the sign of a winning zero is the first occurrence's."""
    from pcgol_b200.dist import decode_minmax_words

    xyz = np.array([[np.nan, 1.0, 0.0], [2.0, np.nan, -0.0], [-3.0, 5.0, 0.5], [4.0, -2.0, 0.0]], np.float32)
    words = _NumpyVgShard(xyz, 0.1, 8).minmax_packed(0).numpy()
    mm = decode_minmax_words(words)
    assert np.isnan(mm[0]) and np.isnan(mm[3])                 # x: point 0 is NaN -> stays NaN
    assert mm[1] == -2.0 and mm[4] == 5.0                       # y: the NaN of point 1 takes no part
    assert mm[2] == 0.0 and not np.signbit(mm[2])               # z min: +0 of point 0 comes first
    assert mm[5] == 0.5
    # the same points as two slices whose second starts the zero run with -0: reduced with MIN like the all-reduce
    a = _NumpyVgShard(xyz[2:3], 0.1, 8).minmax_packed(0).numpy()
    b = _NumpyVgShard(xyz[1:2], 0.1, 8).minmax_packed(1).numpy()
    mm = decode_minmax_words(np.minimum(a, b))
    assert mm[2] == 0.0 and np.signbit(mm[2]) and mm[5] == 0.5  # z: min is the -0 of global point 1
