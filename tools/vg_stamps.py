"""Phase timeline of the fused VoxelGrid kernel (needs a -DPCG_VG_TIMING build via PCG_LIB)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pcgol_b200 as pg
from pcgol_b200 import synth, _lib
pts = synth.lidar_scan(0, n_az=15625)
d = torch.from_numpy(pts).cuda(); out = torch.empty_like(d)
lf = (C.c_float * 3)(0.05, 0.05, 0.05); ck = (C.c_int64 * 3)(128, 128, 128); off = (C.c_int64 * 3)(0, 4, 8); n = C.c_int64(0)
names = ["start", "minmax local", "sync0", "params", "keys"]
for p in range(4): names += [f"p{p} ranked", f"p{p} syncA", f"p{p} walked", f"p{p} scattered", f"p{p} syncB"]
names += ["sort done", "staged+heads", "syncC", "reduced"]
acc = None
for it in range(6):
    _lib.check(_lib.lib.pcg_voxelgrid_filter_dev(d.data_ptr(), len(pts), 12, off, lf, ck, 0, out.data_ptr(), C.byref(n), None))
    st = (C.c_ulonglong * 640)(); k = _lib.lib.pcg_debug_vg_stamps(st)
    t = np.array(st[:k], np.float64)
    print("  last block ends", (st[63] - st[0]) / 1e3, "us after block 0 started; block 0 ended at", (st[k - 1] - st[0]) / 1e3)
    tm = np.array(st[64:64 + k], np.float64)
    if it >= 2:
        acc = (t - t[0]) if acc is None else acc + (t - t[0])
        accm = (tm - t[0]) if it == 2 else accm + (tm - t[0])
acc /= 4; accm /= 4
prev = 0.0
for nm, v, vm in zip(names, acc, accm):
    print(f"{nm:16s} {v / 1e3:8.2f} us  (+{(v - prev) / 1e3:6.2f})   latest CTA {vm / 1e3:8.2f}"); prev = v

tl = np.array(st[128:608], np.float64).reshape(3, 160)
dur = (tl[1] - tl[0]) / 1e3
order = np.argsort(-dur)[:8]
print("slowest tiles in the reduce phase (tile, us, voxel heads):", [(int(t), round(float(dur[t]), 1), int(tl[2][t])) for t in order])
print("median tile:", round(float(np.median(dur[:123])), 1), "us; heads median", int(np.median(tl[2][:123])))
