"""Multi-GPU paths on real devices: sharded ICP with the NCCL all-reduce (needs >= 2 GPUs; on
one GPU the single-rank loop is still checked against the fused single-GPU Fit)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_sharded_icp_single_rank_matches_fast_fit():
    import torch

    import pcgol_b200 as pg
    from pcgol_b200 import dist as pdist, synth

    base, target = synth.icp_pair(seed=3, n=30000, n_az=500)
    idx = pg.Index(base)
    d_t = torch.from_numpy(target).cuda()
    icp = pg.PointToPointICPGradient(pg.PointToPointEvaluator(pg.NearestPointCorresponder(1.0), mode=pg.FAST))
    p = icp.params()
    status, trans, ev, iters = pdist.sharded_icp_fit(pdist.make_gpu_partial(idx, d_t.data_ptr(), len(target), 1.0), p)
    ftrans, fstat = icp.fit(idx, target)
    assert status == 0 and iters == fstat.num_iteration
    # same float64 partial sums folded in a different order -> identical after rounding to float32,
    # up to rare 1-ulp flips amplified over the iterations
    np.testing.assert_allclose(trans, ftrans, rtol=0, atol=1e-5)
    # device-resident loop, one rank: the same reduction order as the host-driven loop -> the same bits
    stream = torch.cuda.current_stream().cuda_stream
    for poll in (0, 5):
        rstatus, rtrans, rstat = pdist.sharded_icp_fit_device(idx, d_t.data_ptr(), len(target), p, stream=stream,
                                                              poll_every=poll)
        assert rstatus == 0 and rstat.num_iteration == iters
        assert rtrans.tobytes() == trans.tobytes()
    # ErrNotEnoughPairs travels through the resident loop like through Fit (icp.go:51-53)
    far = torch.from_numpy((target + np.float32(500.0)).astype(np.float32)).cuda()
    rstatus, rtrans, rstat = pdist.sharded_icp_fit_device(idx, far.data_ptr(), len(target), p, stream=stream)
    assert rstatus == pg._lib.E_NOT_ENOUGH_PAIRS and rstat.num_iteration == 1


def test_fit_multi_one_device_matches_fast_fit():
    """pcg_icp_fit_multi with a single replica: the persistent kernel without any exchange.  Same float64 partial sums
    as the single-GPU fast Fit, folded in a different order."""
    import pcgol_b200 as pg
    from pcgol_b200 import synth

    base, target = synth.icp_pair(seed=3, n=30000, n_az=500)
    idx = pg.Index(base)
    icp = pg.PointToPointICPGradient(pg.PointToPointEvaluator(pg.NearestPointCorresponder(1.0), mode=pg.FAST))
    ftrans, fstat = icp.fit(idx, target)
    mtrans, mstat = icp.fit_multi([idx], target)
    assert mstat.num_iteration == fstat.num_iteration and mstat.n_pairs == fstat.n_pairs
    np.testing.assert_allclose(mtrans, ftrans, rtol=0, atol=1e-6)
    # ErrNotEnoughPairs travels through the persistent loop like through Fit (icp.go:51-53)
    with pytest.raises(pg.ErrNotEnoughPairs) as ei:
        icp.fit_multi([idx], (target + np.float32(500.0)).astype(np.float32))
    assert ei.value.stat.num_iteration == 1
    # strict order cannot be sharded: the entry point always runs the fast mode, whatever the evaluator says
    strict = pg.PointToPointICPGradient(pg.PointToPointEvaluator(pg.NearestPointCorresponder(1.0)))
    strans, _ = strict.fit_multi([idx], target)
    assert strans.tobytes() == mtrans.tobytes()


def test_fit_multi_peer_exchange_matches_single_gpu():
    """One process, several GPUs: index replicas by peer copy, target split, ten float64 sums exchanged over NVLink
    inside the per-device kernels.  Transform within 1e-6 of the single-GPU fast Fit and of the float64 oracle's."""
    import torch

    import pcgol_b200 as pg
    from oracle import oracle as orc
    from pcgol_b200 import synth

    ndev = min(torch.cuda.device_count(), 8)
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    base, target = synth.icp_pair(seed=4, n=60000, n_az=900)
    idx = pg.Index(base)
    replicas = [idx] + [idx.replicate(d) for d in range(1, ndev)]
    q = synth.nn_queries(base, 20000, seed=5)
    ids0, dsq0 = idx.nearest_batch(q, 1.0)
    for r in replicas[1:]:  # a replica answers exactly like the source
        ids, dsq = r.nearest_batch(q, 1.0)
        assert np.array_equal(ids, ids0) and dsq.tobytes() == dsq0.tobytes()
    icp = pg.PointToPointICPGradient(pg.PointToPointEvaluator(pg.NearestPointCorresponder(1.0), mode=pg.FAST))
    ftrans, fstat = icp.fit(idx, target)
    for k in sorted({2, ndev}):
        mtrans, mstat = icp.fit_multi(replicas[:k], target)
        assert mstat.num_iteration == fstat.num_iteration and mstat.n_pairs == fstat.n_pairs
        np.testing.assert_allclose(mtrans, ftrans, rtol=0, atol=1e-6)
    rc, etrans, _, eit = orc.icp_fit(orc.Search(base, "kdtree"), target, orc.icp_params(1.0, f64_accumulate=True))
    assert rc == orc.OK
    np.testing.assert_allclose(mtrans, etrans, rtol=0, atol=1e-5)
    for r in replicas[1:]:
        r.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    import pcgol_b200 as pg
    from pcgol_b200 import dist as pdist, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        base, target = synth.icp_pair(seed=3, n=30000, n_az=500)
        idx = pg.Index(base, device=rank)  # replicated index
        lo, hi = pdist.shard_bounds(len(target), rank, world)
        d_t = torch.from_numpy(target[lo:hi].copy()).cuda()
        icp = pg.PointToPointICPGradient(pg.PointToPointEvaluator(pg.NearestPointCorresponder(1.0), mode=pg.FAST))
        status, trans, ev, iters = pdist.sharded_icp_fit(
            pdist.make_gpu_partial(idx, d_t.data_ptr(), hi - lo, 1.0), icp.params())
        # the same Fit with the loop resident on the device (no host round trip per iteration)
        stream = torch.cuda.current_stream().cuda_stream
        rstatus, rtrans, rstat = pdist.sharded_icp_fit_device(idx, d_t.data_ptr(), hi - lo, icp.params(), stream=stream)
        assert rstatus == status and rstat.num_iteration == iters
        assert rtrans.tobytes() == trans.tobytes(), "resident loop differs from the host-driven loop"
        # one large VoxelGrid: cloud replicated, chunk ranges per rank, all-gather of the counts
        scan = synth.lidar_scan(3, n_az=2000)
        d_scan = torch.from_numpy(scan).cuda()
        d_vg = torch.empty(len(scan) * 12, dtype=torch.uint8, device="cuda")
        m, counts, _ = pdist.sharded_voxelgrid(d_scan.data_ptr(), len(scan), (0.1, 0.1, 0.1), (32, 32, 32), rank, world,
                                               d_vg.data_ptr(), device=rank)
        vg_bytes = d_vg[: m * 12].cpu().numpy().tobytes()
        assert counts[rank] == m and len(counts) == world
        # the same Filter with the POINTS sharded: every rank holds a slice, records travel to their chunk's owner
        data, stride, off = synth.with_fields(scan, extra_u32=1)  # xyz + label: whole records must travel
        slo, shi = pdist.shard_bounds(len(scan), rank, world)
        d_slice = torch.from_numpy(data[slo * stride: shi * stride].copy()).cuda()
        d_pv = torch.empty(len(scan) * stride, dtype=torch.uint8, device="cuda")
        shard = pdist.GpuVgShard(d_slice, shi - slo, stride, off, (0.1, 0.1, 0.1), (32, 32, 32), rank)
        pm, pcounts, _, _, _ = pdist.sharded_voxelgrid_points(shard, slo, rank, world, out=d_pv, n_total=len(scan))
        pv_bytes = d_pv[: pm * stride].cpu().numpy().tobytes()
        assert pcounts[rank] == pm
        # query sharding: disjoint slices of the same queries, index replicated, no collective
        q_all = synth.nn_queries(base, 50000, seed=5)
        qlo, qhi = pdist.shard_bounds(len(q_all), rank, world)
        ids, dsq = idx.nearest_batch(q_all[qlo:qhi], 1.0)
        q.put((rank, status, trans.tobytes(), iters, qlo, ids.tobytes(), dsq.tobytes(), vg_bytes, counts, pv_bytes, pcounts))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_icp_and_queries_world2_nccl():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    import pcgol_b200 as pg
    from pcgol_b200 import synth

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=500) for _ in procs)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert res[0][1] == res[1][1] == 0
    assert res[0][2] == res[1][2] and res[0][3] == res[1][3]  # identical transform on both ranks
    base, target = synth.icp_pair(seed=3, n=30000, n_az=500)
    idx = pg.Index(base)
    icp = pg.PointToPointICPGradient(pg.PointToPointEvaluator(pg.NearestPointCorresponder(1.0), mode=pg.FAST))
    ftrans, fstat = icp.fit(idx, target)
    assert fstat.num_iteration == res[0][3]
    np.testing.assert_allclose(np.frombuffer(res[0][2], np.float32), ftrans, rtol=0, atol=1e-5)
    q_all = synth.nn_queries(base, 50000, seed=5)
    ids, dsq = idx.nearest_batch(q_all, 1.0)
    got_ids = np.concatenate([np.frombuffer(r[5], np.int64) for r in res])
    got_dsq = np.concatenate([np.frombuffer(r[6], np.float32) for r in res])
    assert np.array_equal(got_ids, ids) and got_dsq.tobytes() == dsq.tobytes()
    # sharded VoxelGrid: rank outputs in rank order == the single-GPU Filter, counts all-gathered
    scan = synth.lidar_scan(3, n_az=2000)
    full = pg.VoxelGrid((0.1, 0.1, 0.1), (32, 32, 32)).filter(pg.PointCloud.from_xyz(scan))
    assert b"".join(r[7] for r in res) == full.data[: full.points * 12].tobytes()
    assert res[0][8] == res[1][8] and sum(res[0][8]) == full.points
    # point-sharded Filter of the labelled cloud: rank outputs in rank order == the single-GPU Filter == the oracle
    from oracle import oracle as orc
    data, stride, off = synth.with_fields(scan, extra_u32=1)
    rc, exp = orc.voxelgrid_filter(data, stride, off, (0.1, 0.1, 0.1), (32, 32, 32), mode="sparse")
    assert rc == orc.OK and b"".join(r[9] for r in res) == exp.tobytes()
    assert res[0][10] == res[1][10] and sum(res[0][10]) * stride == len(exp)


def test_point_sharded_voxelgrid_single_rank_equals_filter():
    """World 1: the sharded driver with explicit bounds must reproduce the plain Filter (no collective involved)."""
    import torch

    import pcgol_b200 as pg
    from pcgol_b200 import dist as pdist, synth

    scan = synth.lidar_scan(4, n_az=900)
    for chunk in ((32, 32, 32), (0, 0, 0)):
        data, stride, off = synth.with_fields(scan, extra_u32=1)
        d = torch.from_numpy(data.copy()).cuda()
        out = torch.empty(len(data), dtype=torch.uint8, device="cuda")
        shard = pdist.GpuVgShard(d, len(scan), stride, off, (0.1, 0.1, 0.1), chunk, 0)
        m, counts, _, _, _ = pdist.sharded_voxelgrid_points(shard, 0, 0, 1, out=out, n_total=len(scan))
        hdr = pg.PointCloudHeader(fields=["x", "y", "z", "label"], size=[4] * 4, type=["F", "F", "F", "U"],
                                  count=[1] * 4, width=len(scan))
        full = pg.VoxelGrid((0.1, 0.1, 0.1), chunk).filter(pg.PointCloud(hdr, data))
        assert m == full.points and counts == [m]
        assert out[: m * stride].cpu().numpy().tobytes() == full.data[: m * stride].tobytes()
