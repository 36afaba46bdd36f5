"""Where the time of pcg_icp_fit_multi_dev goes (tuning tool): python tools/multi_probe.py [n_dev]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pcgol_b200 as pg  # noqa: E402
from pcgol_b200 import _lib, dist as pdist, synth  # noqa: E402

ndev = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
scan = synth.lidar_scan(2, n_az=15625)
target = synth.rigid(scan, 2.0, (0.1, 0.1, 0.05), scan.mean(axis=0))
idx = pg.Index(scan)
icp = pg.PointToPointICPGradient(pg.PointToPointEvaluator(pg.NearestPointCorresponder(1.0), mode=pg.FAST))
d_full = torch.from_numpy(target).cuda()
for k in sorted({1, ndev}):
    replicas = [idx] + [idx.replicate(d) for d in range(1, k)]
    slices = []
    for r in range(k):
        a, b = pdist.shard_bounds(len(target), r, k)
        slices.append(torch.from_numpy(np.ascontiguousarray(target[a:b])).to(torch.device("cuda", r)))
    torch.cuda.set_device(0)
    call = lambda: icp.fit_multi_dev(replicas, [t.data_ptr() for t in slices], [len(t) for t in slices])  # noqa: E731
    for _ in range(3):
        call()
    t0 = time.perf_counter()
    for _ in range(5):
        tr, st = call()
    dt = (time.perf_counter() - t0) / 5
    _lib.profile_enable(True)
    for _ in range(3):
        call()
    rep = _lib.profile_report()
    _lib.profile_enable(False)
    print(f"fit_multi_dev n_dev={k}: {dt * 1e3:.3f} ms/alignment, {st.num_iteration} iterations; kernels (us/launch):",
          {n: round(v["total_ms"] * 1e3 / v["launches"], 1) for n, v in rep.items()}, flush=True)
stream = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    icp.fit_dev(idx, d_full.data_ptr(), len(target), stream)
t0 = time.perf_counter()
for _ in range(5):
    icp.fit_dev(idx, d_full.data_ptr(), len(target), stream)
print(f"single-GPU fit_dev: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms/alignment")
