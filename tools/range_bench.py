"""Tuning aid: kernel breakdown of batched Range (1M-pt scan, 200k queries, r = 0.2 m)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcgol_b200 as pg
from pcgol_b200 import synth
tgt = synth.lidar_scan(2, n_az=15625)
rng = np.random.default_rng(1)
for nq, r in ((200_000, 0.2), (200_000, 0.05), (1_000_000, 0.1)):
    sel = rng.choice(len(tgt), nq, replace=False)
    q = (tgt[sel] + rng.normal(0, 0.03, (nq, 3))).astype(np.float32)
    idx = pg.Index(tgt)
    r_f = float(r)
    idx.range_batch(q[:1000], r)
    pg._lib.profile_enable(True)
    t0 = time.perf_counter()
    off, ids, dsq = idx.range_batch(q, r)
    dt = time.perf_counter() - t0
    pg._lib.profile_enable(False)
    import ctypes as C
    r = C.c_void_p()
    offs = (C.c_int64 * 3)(0, 4, 8)
    t0 = time.perf_counter()
    pg._lib.check(
                  pg._lib.lib.pcg_index_range(idx._h, q.ctypes.data, nq, 12, offs, r_f, C.byref(r)))
    dt_c = time.perf_counter() - t0
    pg._lib.lib.pcg_range_free(r)
    print(f"    raw one-call C form: {dt_c*1e3:.1f} ms")
    import torch
    offs_np = torch.zeros(nq + 1, dtype=torch.int64).pin_memory()
    qp = torch.from_numpy(q).pin_memory()
    t0 = time.perf_counter()
    pg._lib.check(pg._lib.lib.pcg_index_range_count(idx._h, qp.data_ptr(), nq, 12, offs, r_f, offs_np.data_ptr()))
    dt1 = time.perf_counter() - t0
    total = int(offs_np[-1])
    outp = torch.empty(total * 16, dtype=torch.uint8).pin_memory()
    for rep in range(2):
        t0 = time.perf_counter()
        pg._lib.check(pg._lib.lib.pcg_index_range_fill(idx._h, qp.data_ptr(), nq, 12, offs, r_f, offs_np.data_ptr(), outp.data_ptr()))
        dt2 = time.perf_counter() - t0
    print(f"    two-call form, pinned caller buffers: count {dt1*1e3:.1f} ms + fill {dt2*1e3:.1f} ms -> {total/(dt1+dt2)/1e6:.0f} M neighbours/s, {nq/(dt1+dt2)/1e6:.2f} M queries/s")
    rep = pg._lib.profile_report()
    tot = sum(v["total_ms"] for v in rep.values())
    print(f"nq {nq} r {r}: neighbours {off[-1]} (avg {off[-1]/nq:.1f}, max {np.diff(off).max()}) e2e {dt*1e3:.1f} ms, kernels {tot:.2f} ms")
    for k, v in sorted(rep.items(), key=lambda x: -x[1]["total_ms"]):
        print(f"    {k:28s} {v['launches']:3d} x {v['total_ms']/v['launches']*1e3:10.1f} us")
