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
