"""Edge cases of the hot path on the GPU, each checked against the oracle: ragged and degenerate
sizes, duplicates, zero ranges, non-finite parameters, tile-boundary sizes of the fused kernel."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.fixture(scope="module")
def pg():
    import pcgol_b200

    return pcgol_b200


def _vg_both(pg, oracle, xyz, leaf, chunk):
    rc, exp = oracle.voxelgrid_filter(xyz.view(np.uint8).reshape(-1), 12, (0, 4, 8), leaf, chunk, mode="sparse")
    from pcgol_b200 import _lib
    out = np.empty(max(1, xyz.size * 4), np.uint8)
    n_out = C.c_int64(0)
    flat = np.ascontiguousarray(xyz).view(np.uint8).reshape(-1)
    grc = _lib.lib.pcg_voxelgrid_filter(flat.ctypes.data if len(flat) else None, len(xyz), 12, (C.c_int64 * 3)(0, 4, 8),
                                        np.asarray(leaf, f32).ctypes.data, np.asarray(chunk, np.int64).ctypes.data, 0,
                                        out.ctypes.data, C.byref(n_out))
    return rc, exp, grc, out[: n_out.value * 12]


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 511, 512, 513, 2047, 2048, 2049, 8191, 8192, 8193, 303103, 303104,
                               303105])
def test_voxelgrid_sizes_around_tile_boundaries(pg, oracle, n):
    rng = np.random.default_rng(n)
    xyz = (rng.random((n, 3), dtype=f32) * np.array([6.0, 5.0, 2.0], f32)).astype(f32)
    xyz -= xyz.min(axis=0)
    for chunk in ((0, 0, 0), (16, 16, 16)):
        rc, exp, grc, got = _vg_both(pg, oracle, xyz, (0.25, 0.25, 0.25), chunk)
        assert rc == oracle.OK and grc == 0
        assert got.tobytes() == exp.tobytes(), (n, chunk)


def test_voxelgrid_all_points_identical_and_duplicates(pg, oracle):
    xyz = np.tile(np.array([[1.5, 2.5, 0.5]], f32), (5000, 1))
    for chunk in ((0, 0, 0), (8, 8, 8)):
        rc, exp, grc, got = _vg_both(pg, oracle, xyz, (0.1, 0.1, 0.1), chunk)
        assert rc == oracle.OK and grc == 0 and got.tobytes() == exp.tobytes()
        assert len(got) == 12
    rng = np.random.default_rng(1)
    base = (rng.random((300, 3), dtype=f32) * f32(3)).astype(f32)
    xyz = base[rng.integers(0, 300, 20000)]  # heavy duplication: long voxels crossing tiles
    xyz = (xyz - xyz.min(axis=0)).astype(f32)
    rc, exp, grc, got = _vg_both(pg, oracle, xyz, (0.5, 0.5, 0.5), (0, 0, 0))
    assert rc == oracle.OK and grc == 0 and got.tobytes() == exp.tobytes()


def test_voxelgrid_one_giant_voxel_spans_many_tiles(pg, oracle):
    rng = np.random.default_rng(2)
    xyz = (rng.random((100000, 3), dtype=f32) * f32(0.9)).astype(f32)
    xyz[0] = 0
    rc, exp, grc, got = _vg_both(pg, oracle, xyz, (1.0, 1.0, 1.0), (0, 0, 0))
    assert rc == oracle.OK and grc == 0
    assert len(exp) == 12 and got.tobytes() == exp.tobytes()  # 100k-term sequential float32 sum, bit-exact


def test_voxelgrid_signed_zero_minimum(pg, oracle):
    # MinMaxVec3 keeps the first occurrence among equal values (-0 == +0): the sign of vMin survives
    xyz = np.array([[0.0, -0.0, 0.0], [-0.0, 0.0, -0.0], [1.0, 1.0, 1.0], [1.01, 1.0, 1.0], [-0.0, -0.0, 0.0]], f32)
    for order in (slice(None), slice(None, None, -1)):
        pts = np.ascontiguousarray(xyz[order])
        rc, exp, grc, got = _vg_both(pg, oracle, pts, (0.5, 0.5, 0.5), (4, 4, 4))
        assert rc == oracle.OK and grc == 0 and got.tobytes() == exp.tobytes()


def test_voxelgrid_bad_leaf_is_reported_not_crashed(pg, oracle):
    from pcgol_b200 import _lib
    xyz = np.array([[0, 0, 0], [1, 1, 1]], f32)
    rc, exp, grc, got = _vg_both(pg, oracle, xyz, (0.0, 0.1, 0.1), (0, 0, 0))  # division by zero -> int(+Inf)
    assert rc == oracle.E_REF_UNDEFINED and grc == _lib.E_REF_UNDEFINED
    rc, exp, grc, got = _vg_both(pg, oracle, xyz, (0.1, 0.1, 0.1), (0, 0, 0))
    assert rc == oracle.OK and grc == 0 and got.tobytes() == exp.tobytes()


def test_nearest_duplicates_zero_range_and_self_queries(pg, oracle):
    rng = np.random.default_rng(3)
    pts = (rng.random((4000, 3), dtype=f32) * f32(4)).astype(f32)
    pts = np.concatenate([pts, pts[:1000], pts[:10]])  # duplicates: lowest index must win
    idx = pg.Index(pts)
    nv = oracle.Search(pts, "naive")
    for mr in (0.0, 1e-30, 0.2, 50.0):
        ids, d = idx.nearest_batch(pts, mr)
        eids, ed = nv.nearest(pts, mr)
        assert np.array_equal(ids, eids) and d.tobytes() == ed.tobytes(), mr
    off, rids, rd = idx.range_batch(pts[:300], 0.0)
    assert off[-1] == 0
    off, rids, rd = idx.range_batch(pts[:300], 0.3)
    eoff, erids, erd = nv.range(pts[:300], 0.3)
    assert np.array_equal(off, eoff) and np.array_equal(rids, erids) and rd.tobytes() == erd.tobytes()


def test_nearest_collinear_and_coplanar_clouds(pg, oracle):
    t = np.linspace(0, 10, 3000, dtype=f32)
    line = np.stack([t, np.zeros_like(t), np.zeros_like(t)], 1)
    g = np.linspace(0, 5, 60, dtype=f32)
    plane = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    plane = np.concatenate([plane, np.full((len(plane), 1), 2.0, f32)], 1).astype(f32)
    rng = np.random.default_rng(4)
    for pts in (line, plane):
        q = (pts[rng.integers(0, len(pts), 2000)] + rng.normal(0, 0.2, (2000, 3))).astype(f32)
        ids, d = pg.Index(pts).nearest_batch(q, 1.0)
        eids, ed = oracle.Search(pts, "naive").nearest(q, 1.0)
        assert np.array_equal(ids, eids) and d.tobytes() == ed.tobytes()


def test_icp_single_iteration_budget_and_custom_updater(pg, oracle):
    from pcgol_b200 import synth
    base, target = synth.icp_pair(seed=9, n=5000, n_az=200)
    idx = pg.Index(base)
    for kw in (dict(max_iteration=1), dict(max_iteration=7, weight=(0.1, 0.2, 0.3, 0.05, 0.05, 0.4)),
               dict(threshold=(1e-4,) * 6, max_iteration=35), dict(threshold=(10.0,) * 6)):
        uf = pg.GradientDescentUpdaterFactory(weight=kw.get("weight", (0,) * 6), threshold=kw.get("threshold", (0,) * 6),
                                              max_iteration=kw.get("max_iteration", 0))
        icp = pg.PointToPointICPGradient(pg.PointToPointEvaluator(pg.NearestPointCorresponder(1.0)), uf)
        trans, stat = icp.fit(idx, target)
        rc, etrans, eev, eit = oracle.icp_fit(oracle.Search(base, "kdtree"), target,
                                              oracle.icp_params(1.0, 0, kw.get("weight"), kw.get("threshold"),
                                                                kw.get("max_iteration", 0)))
        assert rc == oracle.OK and stat.num_iteration == eit, kw
        assert trans.tobytes() == etrans.tobytes(), kw


def test_icp_large_rotation_step_uses_trig_branch(pg, oracle):
    # |delta omega| >= 0.1 rad reaches the sin/cos branch of rodriguesToRotation (rodrigues.go:26-31)
    from pcgol_b200 import synth
    base, _ = synth.icp_pair(seed=4, n=4000, n_az=200)
    target = synth.rigid(base, 25.0, (0.0, 0.0, 0.0), base.mean(axis=0))
    uf = pg.GradientDescentUpdaterFactory(weight=(0.3, 0.3, 0.3, 3.0, 3.0, 3.0), max_iteration=6)
    icp = pg.PointToPointICPGradient(pg.PointToPointEvaluator(pg.NearestPointCorresponder(3.0)), uf)
    trans, stat = icp.fit(pg.Index(base), target)
    rc, etrans, _, eit = oracle.icp_fit(oracle.Search(base, "kdtree"), target,
                                        oracle.icp_params(3.0, 0, (0.3, 0.3, 0.3, 3.0, 3.0, 3.0), None, 6))
    assert rc == oracle.OK and stat.num_iteration == eit
    # float64 sin/cos of the device vs glibc: equal after rounding to float32 except on rare last-ulp flips
    np.testing.assert_allclose(trans, etrans, rtol=0, atol=2e-6)


def _seq_sum(pg, x, exact):
    x = np.ascontiguousarray(x, f32)
    out = (C.c_float * 4)()
    pg._lib.check(pg._lib.lib.pcg_debug_sequential_sum_f32(x.ctypes.data if len(x) else None, len(x), 0, 1 if exact else 0,
                                                          out))
    return f32(out[0])


def test_strict_replay_is_the_sequential_float32_sum(pg):
    """The binade-parallel replay must equal the one-by-one float32 accumulation for ANY stream
    (np.cumsum in float32 is that accumulation), not just for well-behaved residuals."""
    rng = np.random.default_rng(7)
    streams = {
        "normal": rng.normal(0, 1, 100_003).astype(f32),
        "positive": (rng.random(200_000, dtype=f32) * f32(0.3)),
        "tiny_then_huge": np.concatenate([rng.random(5000, dtype=f32) * f32(1e-6), rng.random(5000, dtype=f32) * f32(1e6),
                                          rng.normal(0, 1e-3, 5000).astype(f32)]),
        "cancel": np.concatenate([np.full(3000, 0.1, f32), np.full(3000, -0.1, f32), rng.normal(0, 1, 3000).astype(f32)]),
        "ties_quarters": (rng.integers(-8, 9, 50_000).astype(f32) * f32(0.25)),   # exact ties everywhere
        "ties_halfulp": np.concatenate([[f32(1.0)], np.full(20_000, f32(2.0 ** -24), f32)]),
        "zeros": np.zeros(10_000, f32),
        "sparse": np.where(rng.random(100_000) < 0.01, rng.normal(0, 5, 100_000), 0).astype(f32),
        "wide": (rng.normal(0, 1, 60_000) * np.exp(rng.normal(0, 12, 60_000))).astype(f32),
        "denormal": (rng.random(4000, dtype=f32) * f32(1e-41)),
        "with_inf": np.concatenate([rng.normal(0, 1, 1000).astype(f32), [f32(np.inf)], rng.normal(0, 1, 1000).astype(f32)]),
        "with_nan": np.concatenate([rng.normal(0, 1, 700).astype(f32), [f32(np.nan)], rng.normal(0, 1, 300).astype(f32)]),
        "overflow": np.full(5000, f32(3e38), f32),
        "ramp": np.arange(1, 70_000, dtype=f32),
        "alternating": (np.where(np.arange(90_000) % 2 == 0, 1.0, -1.0) * rng.random(90_000)).astype(f32) * f32(100),
    }
    for n in (0, 1, 2, 255, 256, 257, 511, 512, 513, 1023, 1025):
        streams[f"len{n}"] = rng.normal(0, 3, n).astype(f32)
    for name, x in streams.items():
        with np.errstate(all="ignore"):
            exp = np.cumsum(x, dtype=f32)[-1] if len(x) else f32(0)
        got_seq = _seq_sum(pg, x, exact=False)
        got = _seq_sum(pg, x, exact=True)
        assert got_seq.tobytes() == f32(exp).tobytes() or (np.isnan(exp) and np.isnan(got_seq)), name
        assert got.tobytes() == got_seq.tobytes() or (np.isnan(got) and np.isnan(got_seq)), (name, got, got_seq)


@pytest.mark.gpu
@pytest.mark.parametrize("chunk", [(0, 0, 0), (128, 128, 128), (4096, 4096, 4096)])
def test_voxelgrid_keys_wider_than_32_bits(pg, oracle, chunk):
    # leaf 1 mm over a 4 x 3 x 1.5 m box: 1.8e10 voxels un-chunked (35-bit keys), 9216 chunks x 22-bit keys chunked;
    # the 64-bit sort-key path of the plain Filter and of the sharded one (3 ranks emulated on one GPU)
    import torch
    from pcgol_b200 import dist as pdist

    rng = np.random.default_rng(77)
    n = 30000
    xyz = (rng.random((n, 3), dtype=f32) * np.array([4.0, 3.0, 1.5], f32)).astype(f32)
    xyz[: n // 2] = (xyz[: n // 2] * f32(0.01)).astype(f32)  # half of the points share voxels near the origin
    xyz -= xyz.min(axis=0)
    leaf = (0.001, 0.001, 0.001)
    rc, exp, grc, got = _vg_both(pg, oracle, xyz, leaf, chunk)
    assert rc == oracle.OK and grc == 0
    assert got.tobytes() == exp.tobytes()
    assert len(exp) < n * 12  # some voxels do hold several points
    d_in = torch.from_numpy(xyz).cuda()
    d_out = torch.empty(n * 12, dtype=torch.uint8, device="cuda")
    parts = []
    for rank in range(3):
        m, _, _ = pdist.sharded_voxelgrid(d_in.data_ptr(), n, leaf, chunk, rank, 3, d_out.data_ptr())
        parts.append(d_out[: m * 12].cpu().numpy().copy())
    assert np.concatenate(parts).tobytes() == exp.tobytes()


@pytest.mark.gpu
def test_one_handle_many_threads(pg):
    # the reference's KDTree.Nearest / Range are goroutine-safe (sync.Pool, kdtree.go:44-50; CI runs -race): eight OS
    # threads share one index handle and also run Filters and Fits side by side; every result must equal the
    # single-threaded one bit for bit
    import threading

    from pcgol_b200 import synth

    scan = synth.lidar_scan(5, n_az=600)  # ~38k points
    rng = np.random.default_rng(5)
    idx = pg.Index(scan)
    queries = [(scan[rng.choice(len(scan), 4000)] + rng.normal(0, 0.1, (4000, 3))).astype(f32) for _ in range(8)]
    vg = pg.VoxelGrid((0.1, 0.1, 0.1), (32, 32, 32))
    cloud = pg.PointCloud.from_xyz(scan)
    c, s = np.cos(0.02), np.sin(0.02)
    target = (scan[::3] @ np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], f32).T + f32(0.05)).astype(f32)
    icp = pg.PointToPointICPGradient(pg.PointToPointEvaluator(pg.NearestPointCorresponder(1.0)))

    def work(k):
        ids, dsq = idx.nearest_batch(queries[k], 0.5)
        off, rids, rdsq = idx.range_batch(queries[k][:500], 0.3)
        out = vg.filter(cloud).data.tobytes()
        trans, stat = icp.fit(idx, target)
        return ids.tobytes(), dsq.tobytes(), off.tobytes(), rids.tobytes(), rdsq.tobytes(), out, trans.tobytes(), \
            stat.num_iteration

    expected = [work(k) for k in range(8)]
    got = [None] * 8
    errors = []

    def run(k):
        try:
            for _ in range(3):
                got[k] = work(k)
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=run, args=(k,)) for k in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k in range(8):
        assert got[k] == expected[k], k
