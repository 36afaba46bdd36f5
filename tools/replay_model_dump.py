"""Tuning aid (CPU only): writes the nine per-target ICP term streams of config 1 (evaluator.go:130-144) at the first
and at the last iteration of the oracle's Fit to /tmp/rpm/iter{0,N}.bin ([9][n] float32) for tools/replay_model.cpp."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcgol_b200 import synth
from oracle import oracle as orc

f32 = np.float32
base, target = synth.icp_pair(seed=1)
search = orc.Search(base, "kdtree")
rc, trans, ev, it = orc.icp_fit(search, target, orc.icp_params(1.0))
print("fit", rc, it, ev)


def terms(tr):
    m4 = tr.reshape(4, 4).T.astype(f32)  # column-major
    pt = (target @ m4[:3, :3].T + m4[:3, 3]).astype(f32)
    ids, dsq = search.nearest(pt, 1.0, threads=8)
    m = ids >= 0
    pb = base[np.where(m, ids, 0)]
    z = lambda a: np.where(m, a, 0).astype(f32)
    d = (pt - pb).astype(f32)
    return np.stack([z(dsq), m.astype(f32), z(d[:, 0]), z(d[:, 1]), z(d[:, 2]),
                     z(f32(pt[:, 2] * pb[:, 1]) - f32(pt[:, 1] * pb[:, 2])),
                     z(f32(pt[:, 0] * pb[:, 2]) - f32(pt[:, 2] * pb[:, 0])),
                     z(f32(pt[:, 1] * pb[:, 0]) - f32(pt[:, 0] * pb[:, 1])),
                     z((pt * pt).sum(1))]).astype(f32)


os.makedirs("/tmp/rpm", exist_ok=True)
ident = np.eye(4, dtype=f32).T.reshape(-1)
terms(ident).tofile("/tmp/rpm/iter0.bin")
terms(trans).tofile("/tmp/rpm/iterN.bin")
print("n", len(target))
