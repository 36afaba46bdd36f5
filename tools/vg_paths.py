"""VoxelGrid pipelines side by side on the GPU (tuning tool): bytes of every path against the LSD pipeline
(and the oracle for the small clouds), then CUDA-event time per kernel.
    python tools/vg_paths.py [--big TX TY] [--no-oracle]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pcgol_b200 as pg  # noqa: E402
from pcgol_b200 import _lib, synth  # noqa: E402

NAMES = {0: "auto", 1: "packed", 2: "pairs"}


def run_dev(vg, d_in, n, stride, off, d_out):
    m = vg.filter_dev(d_in.data_ptr(), n, stride, off, d_out.data_ptr())
    return m, d_out[: m * stride].cpu().numpy().tobytes()


def compare(name, data, stride, off, leaf, chunk, paths=(2, 1, 0), oracle=True):
    n = len(data) // stride
    d_in = torch.from_numpy(np.frombuffer(data, np.uint8).copy()).cuda()
    d_out = torch.empty(max(1, n * stride), dtype=torch.uint8, device="cuda")
    vg = pg.VoxelGrid(leaf, chunk)
    ref = None
    res = {}
    for p in paths:
        _lib.set_vg_path(p)
        try:
            m, b = run_dev(vg, d_in, n, stride, off, d_out)
        except Exception as e:  # noqa: BLE001
            res[NAMES[p]] = f"ERR {type(e).__name__} {e}"
            continue
        if ref is None:
            ref = b
        res[NAMES[p]] = (m, "same" if b == ref else "DIFFERENT")
    if oracle:
        from oracle import oracle as orc
        rc, exp = orc.voxelgrid_filter(np.frombuffer(data, np.uint8), stride, off, leaf, chunk, mode="sparse")
        res["oracle"] = "same" if (rc == orc.OK and exp.tobytes() == ref) else f"DIFFERENT rc={rc}"
    _lib.set_vg_path(0)
    print(f"{name:34s} n={n:9d} {res}", flush=True)


def timing(name, xyz, leaf, chunk, paths, reps=10):
    n = len(xyz)
    d_in = torch.from_numpy(xyz.view(np.uint8).reshape(-1).copy()).cuda()
    d_out = torch.empty(n * 12, dtype=torch.uint8, device="cuda")
    vg = pg.VoxelGrid(leaf, chunk)
    for p in paths:
        _lib.set_vg_path(p)
        for _ in range(3):
            vg.filter_dev(d_in.data_ptr(), n, 12, (0, 4, 8), d_out.data_ptr())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            vg.filter_dev(d_in.data_ptr(), n, 12, (0, 4, 8), d_out.data_ptr())
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / reps
        _lib.profile_enable(True)
        for _ in range(reps):
            vg.filter_dev(d_in.data_ptr(), n, 12, (0, 4, 8), d_out.data_ptr())
        rep = _lib.profile_report()
        _lib.profile_enable(False)
        ks = {k: round(v["total_ms"] * 1e3 / reps, 1) for k, v in rep.items()}
        print(f"{name} {NAMES[p]:9s} wall {wall * 1e6:9.1f} us/call  kernels(us/call) {ks}  sum {sum(ks.values()):.1f}",
              flush=True)
    _lib.set_vg_path(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", type=int, nargs=2, default=None)
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--no-compare", action="store_true")
    a = ap.parse_args()
    leaf = (0.05, 0.05, 0.05)
    if not a.no_compare:
        rng = np.random.default_rng(0)
        for n in (1, 5, 100, 4096, 4097, 20000):
            xyz = rng.uniform(-3, 3, (n, 3)).astype(np.float32)
            compare(f"uniform n={n} chunk 16", xyz.tobytes(), 12, (0, 4, 8), (0.1, 0.1, 0.1), (16, 16, 16), oracle=not a.no_oracle)
        scan = synth.lidar_scan(5, n_az=300)
        data, stride, off = synth.with_fields(scan, extra_u32=1)
        compare("scan 19k + label, chunk 64", data.tobytes(), stride, off, (0.1, 0.1, 0.1), (64, 64, 64), oracle=not a.no_oracle)
        compare("scan 19k xyz, unchunked", scan.tobytes(), 12, (0, 4, 8), (0.1, 0.1, 0.1), (0, 0, 0), oracle=not a.no_oracle)
        scan = synth.lidar_scan(3, n_az=1875)
        compare("scan 120k xyz chunk 128", scan.tobytes(), 12, (0, 4, 8), leaf, (128, 128, 128), oracle=not a.no_oracle)
        sh = scan[np.random.default_rng(1).permutation(len(scan))]
        compare("scan 120k shuffled", sh.tobytes(), 12, (0, 4, 8), leaf, (128, 128, 128), oracle=not a.no_oracle)
        dup = np.repeat(scan[:40], 300, axis=0)  # heavy voxels: 300 copies of 40 points -> overflow -> LSD fallback
        compare("12k points in 40 voxels", dup.tobytes(), 12, (0, 4, 8), leaf, (128, 128, 128), oracle=not a.no_oracle)
    scan1m = synth.lidar_scan(2, n_az=15625)
    if not a.no_compare:
        compare("scan 1M xyz chunk 128", scan1m.tobytes(), 12, (0, 4, 8), leaf, (128, 128, 128), oracle=False)
        compare("scan 1M xyz leaf 1mm (u64 keys)", scan1m.tobytes(), 12, (0, 4, 8), (0.001, 0.001, 0.001), (128, 128, 128),
                paths=(2, 1, 0), oracle=False)
    timing("1M", scan1m, leaf, (128, 128, 128), (0, 1, 2))
    if a.big:
        big = synth.tiled_map(a.big[0], a.big[1])
        if not a.no_compare:
            compare(f"map {len(big)}", big.tobytes(), 12, (0, 4, 8), leaf, (128, 128, 128), paths=(2, 1), oracle=False)
        timing(f"map{len(big) // 1000000}M", big, leaf, (128, 128, 128), (1, 2), reps=5)


if __name__ == "__main__":
    main()
