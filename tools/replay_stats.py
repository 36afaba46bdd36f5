"""Tuning aid: fast/slow chunk counts of the exact replay on ICP-like term streams."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcgol_b200 as pg
from pcgol_b200 import synth
from oracle import oracle as orc
base, target = synth.icp_pair(seed=1)
ids, dsq = orc.Search(base, "kdtree").nearest(target, 1.0, threads=8)
m = ids >= 0
pt = target; pb = base[np.where(m, ids, 0)]
f32 = np.float32
streams = {
 "dsq": np.where(m, dsq, 0).astype(f32), "w": m.astype(f32),
 "gx": np.where(m, pt[:,0]-pb[:,0], 0).astype(f32), "gz": np.where(m, pt[:,2]-pb[:,2], 0).astype(f32),
 "gwx": np.where(m, f32(pt[:,2]*pb[:,1]) - f32(pt[:,1]*pb[:,2]), 0).astype(f32),
 "R": np.where(m, (pt*pt).sum(1), 0).astype(f32),
}
for name, x in streams.items():
    out = (C.c_float * 4)()
    pg._lib.lib.pcg_debug_sequential_sum_f32(x.ctypes.data, len(x), 0, 1, out)
    t0 = time.perf_counter()
    for _ in range(5): pg._lib.lib.pcg_debug_sequential_sum_f32(x.ctypes.data, len(x), 0, 1, out)
    dt = (time.perf_counter() - t0) / 5
    print(f"{name:4s} sum {out[0]:.6g} fast {int(out[1])} slow {int(out[2])} walk cycles {int(out[3])} ({out[3]/1.965e3:.1f} us)  call {dt*1e3:.3f} ms  matched {m.mean():.3f}")
# near convergence the gradient terms are zero-mean: the accumulator wanders through binades
rng = np.random.default_rng(5)
for name, x in {"conv_g": rng.normal(0, 0.02, 100_000).astype(f32), "conv_gw": (rng.normal(0, 0.02, 100_000) * rng.uniform(1, 30, 100_000)).astype(f32)}.items():
    out = (C.c_float * 4)()
    pg._lib.lib.pcg_debug_sequential_sum_f32(x.ctypes.data, len(x), 0, 1, out)
    print(f"{name:7s} sum {out[0]:.6g} fast {int(out[1])} slow {int(out[2])} walk cycles {int(out[3])} ({out[3]/1.965e3:.1f} us)")
