"""BASELINE config 5 scale on one GPU: a 50M-point map through the multi-kernel VoxelGrid path,
batched Range and Nearest against it.  Checked against the oracle where it finishes in seconds
(sparse VoxelGrid restatement, KD-tree on a query sample) and through size-independent properties."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
f32 = np.float32


@pytest.fixture(scope="module")
def big_map():
    from pcgol_b200 import synth

    return synth.tiled_map(10, 5)  # 50 x 1M points


def test_voxelgrid_50m_chunked(big_map, oracle):
    import pcgol_b200 as pg

    assert len(big_map) == 50_000_000
    leaf, chunk = (0.05, 0.05, 0.05), (128, 128, 128)
    out = pg.VoxelGrid(leaf, chunk).filter(pg.PointCloud.from_xyz(big_map))
    # bit-exact against the sparse restatement of the reference algorithm
    rc, exp = oracle.voxelgrid_filter(big_map.view(np.uint8).reshape(-1), 12, (0, 4, 8), leaf, chunk, mode="sparse")
    assert rc == oracle.OK
    assert out.points * 12 == len(exp)
    assert out.data.tobytes() == exp.tobytes()
    # properties: fewer points, every centroid inside the cloud's bounding box
    xyz = out.xyz()
    assert 0 < out.points < len(big_map)
    assert (xyz.min(axis=0) >= big_map.min(axis=0)).all() and (xyz.max(axis=0) <= big_map.max(axis=0)).all()


def test_nearest_and_range_on_50m_map(big_map, oracle):
    import pcgol_b200 as pg

    idx = pg.Index(big_map)
    assert len(idx) == 50_000_000
    rng = np.random.default_rng(0)
    sel = rng.choice(len(big_map), 200_000, replace=False)
    q = (big_map[sel] + rng.normal(0, 0.1, (len(sel), 3))).astype(f32)
    ids, dsq = idx.nearest_batch(q, 0.5)
    # property: the reported neighbour is at the reported distance, and no farther than the source point
    hit = ids >= 0
    d = big_map[ids[hit]] - q[hit]
    dd = ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(f32) + d[:, 2] * d[:, 2]).astype(f32)
    assert dd.tobytes() == dsq[hit].tobytes()
    src = big_map[sel] - q
    sd = ((src[:, 0] * src[:, 0] + src[:, 1] * src[:, 1]).astype(f32) + src[:, 2] * src[:, 2]).astype(f32)
    assert (dsq[hit] <= sd[hit]).all()
    assert hit[sd < f32(0.25)].all()
    # oracle on a window of the map (KD-tree over 50M points would take minutes): points within 2 m of a
    # query subset are the only candidates for maxRange 0.5
    sub = q[:2000]
    lo, hi = sub.min(axis=0) - 1, sub.max(axis=0) + 1
    # use one tile so the window stays small
    tile = (big_map[:, 0] < 80) & (big_map[:, 1] < 50)
    qs = sub[(sub[:, 0] < 78) & (sub[:, 1] < 48) & (sub[:, 0] > 1) & (sub[:, 1] > 1)]
    if len(qs):
        cand = np.flatnonzero(tile)
        kd = oracle.Search(big_map[cand], "kdtree")
        eids, edsq = kd.nearest(qs, 0.5, threads=8)
        gids, gdsq = idx.nearest_batch(qs, 0.5)
        assert gdsq.tobytes() == edsq.tobytes()
        assert np.array_equal(gids[eids >= 0], cand[eids[eids >= 0]])
        # Range: same window, r = 0.2
        off, rids, rdsq = idx.range_batch(qs[:500], 0.2)
        eoff, erids, erdsq = kd.range(qs[:500], 0.2)
        assert np.array_equal(off, eoff)
        assert np.array_equal(rids, cand[erids]) and rdsq.tobytes() == erdsq.tobytes()


@pytest.mark.parametrize("layout", [(20, (4, 8, 12)), (13, (1, 5, 9))])
def test_voxelgrid_packed_words_equal_pairs_at_3m_records(layout, oracle):
    """Above 1.2M points the Filter takes the packed-word pipeline (vg_packed.cuh: several super-tiles, four passes,
    records with payload fields / unaligned records); the (key, index) pairs pipeline must give the same bytes, and
    the sparse restatement of voxelgrid.go:35-187 agrees with both."""
    import ctypes as C

    from pcgol_b200 import _lib
    stride, off = layout
    n = 3_000_000
    rng = np.random.default_rng(stride)
    buf = rng.integers(0, 255, size=(n, stride), dtype=np.uint8)
    xyz = (rng.random((n, 3), dtype=f32) * np.array([60.0, 40.0, 6.0], f32)).astype(f32)
    xyz[: n // 3] = (xyz[: n // 3] * f32(0.05)).astype(f32)  # a dense corner: voxels with many members
    for k in range(3):
        buf[:, off[k]:off[k] + 4] = xyz[:, k:k + 1].copy().view(np.uint8)
    flat = np.ascontiguousarray(buf).reshape(-1)
    leaf, chunk = (0.05, 0.05, 0.05), (128, 128, 128)
    outs = []
    for path in (0, 2):
        _lib.set_vg_path(path)
        try:
            out = np.empty(n * stride, np.uint8)
            n_out = C.c_int64(0)
            rc = _lib.lib.pcg_voxelgrid_filter(flat.ctypes.data, n, stride, (C.c_int64 * 3)(*off),
                                               np.asarray(leaf, f32).ctypes.data, np.asarray(chunk, np.int64).ctypes.data,
                                               0, out.ctypes.data, C.byref(n_out))
            assert rc == 0, _lib.last_error()
            outs.append(out[: n_out.value * stride].tobytes())
        finally:
            _lib.set_vg_path(0)
    assert outs[0] == outs[1]
    rc, exp = oracle.voxelgrid_filter(flat, stride, off, leaf, chunk, mode="sparse")
    assert rc == oracle.OK and exp.tobytes() == outs[0]
