"""Tuning aid (needs build_variants/libpcg_nnstats.so via PCG_LIB): node/leaf visits per query."""
import ctypes as C, numpy as np, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcgol_b200 as pg
from pcgol_b200 import synth
tgt = synth.lidar_scan(2, n_az=15625)
q = synth.nn_queries(tgt, 2_000_000, seed=3)
idx = pg.Index(tgt)
out = (C.c_ulonglong * 4)()
pg._lib.lib.pcg_debug_nn_stats(out)
idx.nearest_batch(q, 1.0)
pg._lib.lib.pcg_debug_nn_stats(out)
nq = out[3]
print(f"queries {nq}  node steps/query {out[0]/nq:.1f}  leaf scans/query {out[1]/nq:.1f}  pushes/query {out[2]/nq:.1f}")
rng = np.random.default_rng(0)
u = (rng.random((1_000_000, 3), dtype=np.float32) * np.array([80, 50, 3], np.float32))
idx2 = pg.Index(u); pg._lib.lib.pcg_debug_nn_stats(out)
idx2.nearest_batch((u[:500000] + rng.normal(0, 0.05, (500000, 3))).astype(np.float32), 1.0)
pg._lib.lib.pcg_debug_nn_stats(out); nq = out[3]
print(f"uniform cloud: node steps/query {out[0]/nq:.1f}  leaf scans/query {out[1]/nq:.1f}  pushes/query {out[2]/nq:.1f}")
