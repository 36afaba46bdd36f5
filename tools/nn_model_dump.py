import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from pcgol_b200 import synth
tgt = synth.lidar_scan(2, n_az=15625)
q = synth.nn_queries(tgt, 10_000_000, seed=3)
rng = np.random.default_rng(0)
sel = np.sort(rng.choice(len(q), 200_000, replace=False))
tgt.tofile('/tmp/nnm/tgt.bin'); q[sel].tofile('/tmp/nnm/q.bin')
print(len(tgt), len(sel), tgt.min(0), tgt.max(0))
