import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pcgol_b200 as pg
from pcgol_b200 import synth
from oracle import oracle as orc
pts = synth.lidar_scan(7)[:100_000].copy()
q = synth.nn_queries(pts, 2_500_001, seed=5)
idx = pg.Index(pts)
ids, dsq = idx.nearest_batch(q, 0.8)
eids, edsq = orc.Search(pts, "kdtree").nearest(q, 0.8, threads=8)
bad = np.flatnonzero((ids != eids) | (dsq.view(np.uint32) != edsq.view(np.uint32)))
print("mismatches", len(bad), bad[:20], "slices", np.unique(bad >> 20, return_counts=True))
for b in bad[:5]:
    print(b, ids[b], dsq[b], eids[b], edsq[b], q[b])
# small batches over the same queries
ids2 = np.concatenate([idx.nearest_batch(q[i:i + 1_000_000], 0.8)[0] for i in range(0, len(q), 1_000_000)])
print("vs unsliced calls:", int((ids2 != ids).sum()), "unsliced vs oracle:", int((ids2 != eids).sum()))
nv = orc.Search(pts, "naive")
nid, nd = nv.nearest(q[bad[:200]], 0.8)
print("naive agrees with gpu on", int((nid == ids[bad[:200]]).sum()), "of", len(nid), "; with kdtree on", int((nid == eids[bad[:200]]).sum()))
