"""Per-kernel time of the index build (pcg_profile_enable): python tools/build_prof.py [n_az]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pcgol_b200 as pg
from pcgol_b200 import synth, _lib
n_az = int(sys.argv[1]) if len(sys.argv) > 1 else 15625
tgt = synth.lidar_scan(2, n_az=n_az)
d = torch.from_numpy(tgt).cuda()
for _ in range(3):
    pg.Index.from_device(d.data_ptr(), len(tgt)).close()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    pg.Index.from_device(d.data_ptr(), len(tgt)).close()
torch.cuda.synchronize()
print(f"n={len(tgt)} build {1e3 * (time.perf_counter() - t0) / 10:.3f} ms (wall, incl. alloc/free)")
_lib.profile_enable(True)
for _ in range(5):
    pg.Index.from_device(d.data_ptr(), len(tgt)).close()
torch.cuda.synchronize()
rep = _lib.profile_report()
_lib.profile_enable(False)
tot = 0
for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["total_ms"]):
    print(f"  {k:40s} {v['launches'] / 5:6.1f} launches  {1e3 * v['total_ms'] / 5:9.1f} us per build")
    tot += v["total_ms"] / 5
print(f"  kernels total {1e3 * tot:.1f} us")
