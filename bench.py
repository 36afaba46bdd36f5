#!/usr/bin/env python
"""bench.py — headline benchmark of the pcgol hot path on B200.

Primary line (BASELINE.json configs[1]): VoxelGrid downsample of a 1M-point synthetic
64-beam scan at 0.05 m leaf, float32 x,y,z records (stride 12), ChunkSize{128,128,128}.
A "step" is one Filter pass over one cloud.  `value` is measured with the cloud resident
in HBM (pcg_voxelgrid_filter_dev); `e2e` goes through the host-buffer C-ABI call
(pcg_voxelgrid_filter) with pinned host input/output, copies inside the timed region.
The same JSON line carries, under "extra", the other two metrics of the path at N=1:
batched Nearest (configs[2]: 10M queries vs a 1M-point target, maxRange 1 m) and
point-to-point ICP (configs[0]: 100k-point scan vs a 5 deg / 0.3 m perturbed copy).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-extra]

N > 1 is launched by torchrun (one rank per GPU); the workload shards by independent
clouds (weak scaling, no data-path collective); timing = max over ranks.
--impl reference times the CPU restatement of the reference (oracle/, the reference is Go and
there is no Go toolchain in the image) with all host threads, on the same config/metric.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LEAF = (0.05, 0.05, 0.05)
CHUNK = (128, 128, 128)
N_AZ_1M = 15625
ROTATE = 16  # distinct device copies of the input cycled between steps: 16 x 12 MB > 126 MB L2
CACHE = os.environ.get("PCGOL_BENCH_CACHE", "/tmp/pcgol_b200_cache")
WORKLOAD = "voxelgrid 1M-pt synthetic 64-beam scan, leaf 0.05 m, xyz f32 stride 12, ChunkSize{128,128,128}"


_REAL_STDOUT = None


def emit(obj):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def cached_scan(seed: int, n_az: int) -> np.ndarray:
    from pcgol_b200 import synth

    os.makedirs(CACHE, exist_ok=True)
    path = os.path.join(CACHE, f"scan_{seed}_{n_az}.npy")
    if os.path.exists(path):
        return np.load(path)
    a = synth.lidar_scan(seed, n_az=n_az)
    tmp = f"{path}.{os.getpid()}.tmp.npy"
    np.save(tmp, a)
    os.replace(tmp, path)
    return a


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU arm ----
def cpu_voxelgrid(scan: np.ndarray, threads: int, repeats: int):
    """Reference algorithm (literal dense-array restatement) on `threads` independent filters."""
    from oracle import oracle as orc

    buf = scan.view(np.uint8).reshape(-1)
    orc.lib()

    def one():
        rc, out = orc.voxelgrid_filter(buf, 12, (0, 4, 8), LEAF, CHUNK, mode="dense")
        assert rc == orc.OK

    t0 = time.perf_counter()
    for _ in range(repeats):
        if threads == 1:
            one()
        else:
            th = [threading.Thread(target=one) for _ in range(threads)]
            [t.start() for t in th]
            [t.join() for t in th]
    dt = time.perf_counter() - t0
    return threads * repeats * len(scan) / dt / 1e6, dt  # Mpts/s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import psutil

    cores = os.cpu_count() or 1
    threads = max(1, min(cores, int(psutil.virtual_memory().available * 0.5 // (200 << 20))))
    scan = cached_scan(2, N_AZ_1M)
    for _ in range(min(args.warmup, 1)):
        cpu_voxelgrid(scan, threads, 1)
    times = []
    for _ in range(args.steps):
        _, dt = cpu_voxelgrid(scan, threads, 1)
        times.append(dt)
        if sum(times) > 150:  # bounded run
            break
    steps = len(times)
    total = sum(times)
    value = threads * steps * len(scan) / total / 1e6
    line = {
        "impl": "reference", "metric": "VoxelGrid Mpts/s", "value": value, "unit": "Mpts/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * total / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "points_per_step": threads * len(scan)},
        "cpu_baseline": {"value": value, "unit": "Mpts/s", "cores": threads, "kind": "port",
                         "sample": f"{threads} threads x one full 1M-pt Filter per step (C++ restatement of "
                                   "voxelgrid.go:35-187, dense voxel array; not Go: no Go toolchain in the image)"},
        "e2e": {"value": value, "unit": "Mpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# --------------------------------------------------------------------------- GPU arm ----
def timed_region(dist, torch, fn, steps):
    """barrier + sync, K steps between CUDA events on the current stream, sync + barrier; max over ranks."""
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    return ms


def wall_region(dist, torch, fn, steps):
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        fn(i)
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    return ms


def profile_kernels(pg, torch, fn, steps):
    """Per-kernel CUDA-event durations over `steps` steps (separate pass, same command)."""
    pg._lib.profile_enable(True)
    try:
        for i in range(steps):
            fn(i)
        torch.cuda.synchronize()
    finally:
        pg._lib.profile_enable(False)
    return pg._lib.profile_report()


def ncu_traffic(kernel_name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
    (profiles/traffic.json, written by tools/ncu_summary.py traffic); None if that kernel was not captured."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        table = json.load(open(path))
    except Exception:
        return None
    for key, val in table.items():
        if key in kernel_name:
            return val.get("dram_bytes_per_launch")
    return None


def dominant(report, algo_bytes, peak):
    """Roofline object for the kernel with the largest share of the step."""
    if not report:
        return None, {}
    total = sum(v["total_ms"] for v in report.values())
    shares = {k: {"launches": v["launches"], "avg_us": 1e3 * v["total_ms"] / v["launches"],
                  "share": v["total_ms"] / total} for k, v in report.items()}
    name = max(report, key=lambda k: report[k]["total_ms"])
    avg_s = report[name]["total_ms"] / report[name]["launches"] / 1e3
    b = algo_bytes.get(name)
    if b is None:  # template arguments are part of the recorded names: match on the kernel's base name
        b = next((v for k, v in algo_bytes.items() if k in name), None)
    roof = {"bound": "hbm", "kernel": name, "avg_launch_us": avg_s * 1e6, "share_of_step": shares[name]["share"],
            "algorithmic_bytes_per_launch": b, "achieved": (b / avg_s / 1e9) if b else None, "peak": peak[0],
            "peak_source": peak[1], "unit": "GB/s", "frac": (b / avg_s / 1e9 / peak[0]) if b else None,
            "traffic": ncu_traffic(name),
            "timing": "CUDA events around every launch of this kernel over K steps (separate pass, same command)"}
    return roof, shares


def bench_voxelgrid(pg, torch, dist, rank, args, peak):
    scan = cached_scan(2 + rank, N_AZ_1M)
    n = len(scan)
    dev = torch.device("cuda")
    d_in = [torch.from_numpy(scan.view(np.uint8).reshape(-1)).to(dev) for _ in range(ROTATE)]
    d_out = torch.empty(n * 12, dtype=torch.uint8, device=dev)
    vg = pg.VoxelGrid(LEAF, CHUNK, device=torch.cuda.current_device())
    stream = torch.cuda.current_stream().cuda_stream
    m_box = [0]

    def step(i):
        m_box[0] = vg.filter_dev(d_in[i % ROTATE].data_ptr(), n, 12, (0, 4, 8), d_out.data_ptr(), stream)

    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:  # one nvidia-smi poller per box is enough; eight of them perturb the host
        sampler.start()
    for i in range(args.warmup):
        step(i)
    l0 = pg.kernel_launch_count()
    ms = timed_region(dist, torch, step, args.steps)
    launches = pg.kernel_launch_count() - l0
    # the timed region lasts a few ms (K steps of ~0.2 ms): keep the same kernels running for
    # ~0.5 s more so that the 100 ms nvidia-smi sampler sees the clocks under this load
    t_hold = time.perf_counter()
    i = 0
    while time.perf_counter() - t_hold < 0.5:
        step(i)
        i += 1
    torch.cuda.synchronize()
    clocks = sampler.stop()
    clocks["window"] = "warm-up + timed region + 0.5 s of the same steps"
    m = m_box[0]

    # end to end: pinned host buffers through the host C-ABI call (synchronous, like Filter.Filter).
    # `callers` host threads each run the same call on their own cloud buffers, the way goroutines would:
    # the library gives every OS thread its own stream, so one caller's PCIe copies overlap another's kernel.
    leaf = np.asarray(LEAF, np.float32)
    chunk = np.asarray(CHUNK, np.int64)
    off = (C.c_int64 * 3)(0, 4, 8)
    device = torch.cuda.current_device()

    def make_caller():
        h_in = torch.from_numpy(scan.view(np.uint8).reshape(-1).copy()).pin_memory()
        h_out = torch.empty(n * 12, dtype=torch.uint8).pin_memory()
        n_out = C.c_int64(0)

        def call():
            rc = pg._lib.lib.pcg_voxelgrid_filter(h_in.data_ptr(), n, 12, off, leaf.ctypes.data, chunk.ctypes.data,
                                                  device, h_out.data_ptr(), C.byref(n_out))
            assert rc == 0, pg._lib.last_error()
            return n_out.value

        return call

    def run_e2e(callers):
        """`callers` persistent host threads; each warms up its own stream/buffers, then all start together."""
        calls = [make_caller() for _ in range(callers)]
        # at least 24 calls per caller: the region is bracketed by Python barriers whose wake-up latency (tenths of a
        # millisecond) must stay small against the calls it times
        per = max(24, args.steps // callers)
        ready = threading.Barrier(callers + 1)
        start = threading.Barrier(callers + 1)
        stop = threading.Barrier(callers + 1)

        def worker(c):
            for _ in range(3):
                assert c() == m
            ready.wait()
            start.wait()
            for _ in range(per):
                c()
            stop.wait()

        th = [threading.Thread(target=worker, args=(c,)) for c in calls]
        [t.start() for t in th]
        ready.wait()  # warm-up (stream creation, pool growth, first copies) stays outside the timed region

        def region(_):
            start.wait()  # releases the warmed-up callers
            stop.wait()   # all calls have returned (each call is synchronous)

        ms_ = wall_region(dist, torch, region, 1)
        [t.join() for t in th]
        return ms_, per * callers

    def run_e2e_median(callers, reps=5):
        """The timed window is a few milliseconds of host-driven calls: thread start-up jitter moves it by tens of
        per cent, so the region is repeated and the median repetition reported (every repetition times `steps` calls)."""
        runs = sorted(run_e2e(callers) for _ in range(reps))
        return runs[len(runs) // 2]

    e2e1_ms, e2e1_steps = run_e2e_median(1)
    # concurrent callers per GPU: up to three, but never more host threads than the box has cores to spare
    cpus = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    E2E_CALLERS = max(1, min(3, cpus // (2 * max(1, args.gpus))))
    e2e_ms, e2e_steps = run_e2e_median(E2E_CALLERS)

    world = args.gpus
    # algorithmic bytes per launch of each kernel of the pipeline
    kb = 4  # (chunk id, voxel key) needs 29 bits for this config -> 32-bit sort keys
    algo = {
        "voxelgrid_fused_kernel<IPT>": 12 * n + 12 * m,  # the whole Filter is this one kernel: N*stride in, M*stride out
        "minmax_kernel": 12 * n,
        "(voxel_key_kernel<K>)": 12 * n + kb * n,
        "(histogram_kernel<K>)": kb * n,
        "(onesweep_kernel<K, IPT>)": 2 * (kb + 4) * n,
        "(voxel_reduce_kernel<K>)": (kb + 4) * n + 12 * n + 12 * m,
    }
    report = profile_kernels(pg, torch, step, args.steps)
    roof, shares = dominant(report, algo, peak)
    step_s = ms / args.steps / 1e3
    pipeline_bytes = 12 * n + 12 * m  # N*stride read + M*stride written
    res = {
        "value": world * n * args.steps / (ms / 1e3) / 1e6,
        "ms_per_step": ms / args.steps,
        "e2e": {"value": world * n * e2e_steps / (e2e_ms / 1e3) / 1e6, "unit": "Mpts/s",
                "h2d_bytes_per_step": 12 * n, "d2h_bytes_per_step": 12 * m, "ms_per_step": e2e_ms / e2e_steps,
                "concurrent_callers": E2E_CALLERS, "steps": e2e_steps,
                "single_caller": {"value": world * n * e2e1_steps / (e2e1_ms / 1e3) / 1e6,
                                  "ms_per_step": e2e1_ms / e2e1_steps},
                "timer": "host wall clock around the synchronous C-ABI calls (pinned host in/out), max over ranks; median of 5 "
                         "repetitions of the K-step region"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "kernels": shares,
        "pipeline_roofline": {"algorithmic_bytes_per_step": pipeline_bytes,
                              "achieved": pipeline_bytes / step_s / 1e9, "peak": peak[0], "unit": "GB/s",
                              "frac": pipeline_bytes / step_s / 1e9 / peak[0]},
        "voxels_out": int(m),
        "scan": scan,
    }
    return res


def bench_nn(pg, torch, dist, rank, args, peak, nq_total=10_000_000):
    from pcgol_b200 import synth

    target = cached_scan(2, N_AZ_1M)
    nq = nq_total
    q = synth.nn_queries(target, nq, seed=3 + rank)
    dev = torch.device("cuda")
    device = torch.cuda.current_device()
    stream = torch.cuda.current_stream().cuda_stream
    d_t = torch.from_numpy(target).to(dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    idx = pg.Index.from_device(d_t.data_ptr(), len(target), device=device, stream=stream)
    torch.cuda.synchronize()
    build_ms_cold = 1e3 * (time.perf_counter() - t0)
    builds = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        tmp = pg.Index.from_device(d_t.data_ptr(), len(target), device=device, stream=stream)
        b.record()
        torch.cuda.synchronize()
        builds.append(a.elapsed_time(b))
        tmp.close()
    d_q = torch.from_numpy(q).to(dev)
    d_ids = torch.empty(nq, dtype=torch.int32, device=dev)
    d_dsq = torch.empty(nq, dtype=torch.float32, device=dev)

    def step(i):
        idx.nearest_dev(d_q.data_ptr(), nq, 1.0, d_ids.data_ptr(), d_dsq.data_ptr(), stream)

    for i in range(args.warmup):
        step(i)
    steps = max(3, min(args.steps, 10))
    l0 = pg.kernel_launch_count()
    ms = timed_region(dist, torch, step, steps)
    launches = pg.kernel_launch_count() - l0

    h_q = torch.from_numpy(q).pin_memory()
    h_out = torch.empty(nq * 16, dtype=torch.uint8).pin_memory()
    off = (C.c_int64 * 3)(0, 4, 8)

    def e2e_step(i):
        rc = pg._lib.lib.pcg_index_nearest(idx._h, h_q.data_ptr(), nq, 12, off, 1.0, h_out.data_ptr())
        assert rc == 0, pg._lib.last_error()

    e2e_step(0)
    e2e_ms = wall_region(dist, torch, e2e_step, 3)
    report = profile_kernels(pg, torch, step, steps)
    algo = {"nearest_simple_kernel": 20 * nq + 16 * len(target), "nearest_kernel": 20 * nq + 16 * len(target)}
    roof, shares = dominant(report, algo, peak)
    ids = d_ids.cpu().numpy()
    dsq = d_dsq.cpu().numpy()
    return {
        "metric": "NN queries/s", "unit": "queries/s",
        "config": {"workload": "batched Nearest: 10M queries (target + N(0,0.3 m) jitter) vs 1M-pt scan, maxRange 1 m",
                   "l2": "inputs+outputs 200 MB > 126 MB L2"},
        "value": args.gpus * nq * steps / (ms / 1e3), "ms_per_step": ms / steps, "steps": steps,
        "e2e": {"value": args.gpus * nq * 3 / (e2e_ms / 1e3), "unit": "queries/s", "h2d_bytes_per_step": 12 * nq,
                "d2h_bytes_per_step": 16 * nq, "ms_per_step": e2e_ms / 3},
        "index_build_ms": {"first_call": build_ms_cold, "warm": float(np.median(builds))},
        "index_bytes": idx.device_bytes(), "gpu_launches": int(launches), "roofline": roof, "kernels": shares,
        "hit_fraction": float((ids >= 0).mean()),
        "_check": (target, q, ids, dsq),
    }


def bench_icp(pg, torch, dist, rank, args, peak):
    from pcgol_b200 import synth

    base, target = synth.icp_pair(seed=1)  # the same pair on every rank: equal units for weak scaling
    dev = torch.device("cuda")
    device = torch.cuda.current_device()
    stream = torch.cuda.current_stream().cuda_stream
    d_b = torch.from_numpy(base).to(dev)
    d_t = torch.from_numpy(target).to(dev)
    idx = pg.Index.from_device(d_b.data_ptr(), len(base), device=device, stream=stream)
    out = {}
    steps = max(3, min(args.steps, 10))
    # strict / fast: the reference's gradient-descent updater (bit-exact / float64 sums).  gauss_newton: NOT the
    # reference's algorithm (normal equations solved per iteration) - reported beside it, never as it.
    for mode_name, mode, factory in (("strict", pg.STRICT, None), ("fast", pg.FAST, None),
                                     ("gauss_newton_fast", pg.FAST, pg.GaussNewtonUpdaterFactory())):
        icp = pg.PointToPointICPGradient(pg.PointToPointEvaluator(pg.NearestPointCorresponder(1.0), mode=mode), factory)
        res = {}

        def step(i):
            res["trans"], res["stat"] = icp.fit_dev(idx, d_t.data_ptr(), len(target), stream)

        for i in range(max(1, args.warmup // 2)):
            step(i)
        l0 = pg.kernel_launch_count()
        ms = timed_region(dist, torch, step, steps)
        launches = pg.kernel_launch_count() - l0
        report = profile_kernels(pg, torch, step, 2)
        iters = res["stat"].num_iteration
        algo = {"(icp_terms_kernel<PCG_ICP_STRICT>)": 12 * len(target) + 36 * len(target),
                "(icp_terms_kernel<PCG_ICP_FAST>)": 12 * len(target),
                "icp_replay_kernel": 36 * len(target), "icp_replay_sums_kernel": 36 * len(target),
                "icp_replay_summaries_kernel": 36 * len(target), "icp_replay_walk_kernel": 36 * len(target) // 16}
        roof, shares = dominant(report, algo, peak)
        h_t = torch.from_numpy(target).pin_memory()
        p = icp.params()
        tr = np.zeros(16, np.float32)
        st = pg._lib.IcpStat()
        off = (C.c_int64 * 3)(0, 4, 8)

        def e2e_step(i):
            rc = pg._lib.lib.pcg_icp_fit(idx._h, h_t.data_ptr(), len(target), 12, off, C.byref(p), tr.ctypes.data,
                                         C.byref(st))
            assert rc == 0, pg._lib.last_error()

        e2e_step(0)
        e2e_ms = wall_region(dist, torch, e2e_step, 3)
        out[mode_name] = {
            "value": args.gpus * steps / (ms / 1e3), "ms_per_alignment": ms / steps, "iterations": int(iters),
            "e2e": {"value": args.gpus * 3 / (e2e_ms / 1e3), "unit": "alignments/s",
                    "h2d_bytes_per_step": 12 * len(target), "d2h_bytes_per_step": 64 + C.sizeof(pg._lib.IcpStat)},
            "gpu_launches": int(launches), "roofline": roof, "kernels": shares,
            "trans": [float(x) for x in res["trans"]], "final_value": float(res["stat"].evaluated.value),
        }
    return {"metric": "ICP alignments/s", "unit": "alignments/s",
            "config": {"workload": "point-to-point ICP Fit (<= 20 iterations, default updater, MaxDist 1 m) of a "
                                   "100k-pt synthetic scan vs a 5 deg / 0.3 m perturbed copy; index prebuilt"},
            "modes": out, "_check": (base, target)}


def bench_icp_farm(pg, torch, dist, rank, args, peak, pairs_per_gpu=64, distinct=4):
    """BASELINE config 4: independent 64-beam scan pairs (~120k pts each, pose perturbed by <= 5 deg / 0.3 m),
    dealt to the ranks; per GPU pcg_icp_fit_pairs_dev builds one index per pair and overlaps the fits on streams.
    `distinct` scan pairs are generated (ray casting on the host is slow) and cycled."""
    from pcgol_b200 import synth

    dev = torch.device("cuda")
    device = torch.cuda.current_device()
    stream = torch.cuda.current_stream().cuda_stream
    host = []
    for k in range(distinct):
        key = f"pair_{k}"  # the same pairs on every rank: equal units for weak scaling
        path = os.path.join(CACHE, key + ".npz")
        if os.path.exists(path):
            z = np.load(path)
            host.append((z["b"], z["t"]))
        else:
            b, t = synth.scan_pair(k)
            os.makedirs(CACHE, exist_ok=True)
            np.savez(path + f".{os.getpid()}.tmp.npz", b=b, t=t)
            os.replace(path + f".{os.getpid()}.tmp.npz", path)
            host.append((b, t))
    d = [(torch.from_numpy(b).to(dev), torch.from_numpy(t).to(dev)) for b, t in host]
    sel = [d[i % distinct] for i in range(pairs_per_gpu)]
    out = {}
    # strict / fast: the reference's gradient-descent updater (bit-exact / float64 sums).  gauss_newton: NOT the
    # reference's algorithm (normal equations solved per iteration) - reported beside it, never as it.
    for mode_name, mode, factory in (("strict", pg.STRICT, None), ("fast", pg.FAST, None),
                                     ("gauss_newton_fast", pg.FAST, pg.GaussNewtonUpdaterFactory())):
        icp = pg.PointToPointICPGradient(pg.PointToPointEvaluator(pg.NearestPointCorresponder(1.0), mode=mode), factory)
        res = {}

        def step(i):
            res["r"] = icp.fit_pairs_dev([b.data_ptr() for b, _ in sel], [len(b) for b, _ in sel],
                                         [t.data_ptr() for _, t in sel], [len(t) for _, t in sel], device=device,
                                         stream=stream)

        step(0)
        step(1)  # two warm-up calls: the stream-ordered pool reaches its steady size for this mode's buffers
        ms = timed_region(dist, torch, step, 2)
        trans, iters, status, _ = res["r"]
        out[mode_name] = {"value": args.gpus * pairs_per_gpu * 2 / (ms / 1e3), "ms_per_pair": ms / 2 / pairs_per_gpu,
                          "iterations_mean": float(np.mean(iters)), "failed": int((status != 0).sum())}
    return {"metric": "ICP scan pairs/s (index build + Fit)", "unit": "pairs/s", "scaling": "weak",
            "config": {"workload": f"{pairs_per_gpu} scan pairs per GPU ({distinct} distinct, ~120k pts each), "
                                   "index built per pair, <= 20 iterations"},
            "modes": out}


def bench_config5(pg, torch, dist, rank, args, peak):
    """BASELINE config 5 on one GPU (opt-in, --config5): 50M-point map, VoxelGrid (multi-kernel path, HBM-sized
    working set) and batched Range (1M queries, r = 0.2 m) against the 50M-point index."""
    from pcgol_b200 import synth

    big = synth.tiled_map(10, 5)
    n = len(big)
    dev = torch.device("cuda")
    device = torch.cuda.current_device()
    stream = torch.cuda.current_stream().cuda_stream
    d_in = torch.from_numpy(big).to(dev)
    d_out = torch.empty(n * 12, dtype=torch.uint8, device=dev)
    vg = pg.VoxelGrid(LEAF, CHUNK, device=device)
    m_box = [0]

    def step(i):
        m_box[0] = vg.filter_dev(d_in.data_ptr(), n, 12, (0, 4, 8), d_out.data_ptr(), stream)

    step(0)
    step(1)
    ms = timed_region(dist, torch, step, 5)
    report = profile_kernels(pg, torch, step, 3)
    m = m_box[0]
    kb = 8
    algo = {"minmax_kernel": 12 * n, "(voxel_key_kernel<K>)": 12 * n + kb * n, "(onesweep_kernel<K, IPT>)": 2 * (kb + 4) * n,
            "(voxel_reduce_kernel<K>)": (kb + 4) * n + 12 * n + 12 * m}
    roof, shares = dominant(report, algo, peak)
    step_s = ms / 5 / 1e3
    out = {"voxelgrid": {"points": n, "voxels_out": int(m), "value_mpts": n / step_s / 1e6, "ms_per_step": ms / 5,
                         "roofline": roof, "kernels": shares,
                         "pipeline_roofline": {"algorithmic_bytes_per_step": 12 * n + 12 * m,
                                               "achieved": (12 * n + 12 * m) / step_s / 1e9, "peak": peak[0],
                                               "frac": (12 * n + 12 * m) / step_s / 1e9 / peak[0]}}}
    t0 = time.perf_counter()
    idx = pg.Index.from_device(d_in.data_ptr(), n, device=device, stream=stream)
    torch.cuda.synchronize()
    out["index_build_ms"] = 1e3 * (time.perf_counter() - t0)
    out["index_bytes"] = idx.device_bytes()
    rng = np.random.default_rng(5)
    sel = rng.choice(n, 1_000_000, replace=False)
    q = (big[sel] + rng.normal(0, 0.05, (len(sel), 3))).astype(np.float32)
    t0 = time.perf_counter()
    off, ids, dsq = idx.range_batch(q, 0.2)
    dt = time.perf_counter() - t0
    out["range"] = {"queries": len(q), "neighbours": int(off[-1]), "queries_per_s_e2e": len(q) / dt,
                    "neighbours_per_s_e2e": int(off[-1]) / dt, "radius": 0.2}
    d_q = torch.from_numpy(q).to(dev)
    d_ids = torch.empty(len(q), dtype=torch.int32, device=dev)
    d_d = torch.empty(len(q), dtype=torch.float32, device=dev)

    def nstep(i):
        idx.nearest_dev(d_q.data_ptr(), len(q), 1.0, d_ids.data_ptr(), d_d.data_ptr(), stream)

    nstep(0)
    ms = timed_region(dist, torch, nstep, 5)
    out["nearest"] = {"queries": len(q), "value": len(q) * 5 / (ms / 1e3), "ms_per_step": ms / 5}
    return out


def bench_voxelgrid_sharded(pg, torch, dist, rank, args, peak):
    """BASELINE config 5 (VoxelGrid part) over the job's GPUs: ONE 50M-point Filter, the cloud replicated, every rank
    filtering a balanced range of chunk ids (dist.sharded_voxelgrid); strong scaling, the step includes the chunk
    histogram, its read-back and the all-gather of the counts."""
    from pcgol_b200 import dist as pdist, synth

    big = synth.tiled_map(10, 5)
    n = len(big)
    dev = torch.device("cuda")
    device = torch.cuda.current_device()
    stream = torch.cuda.current_stream().cuda_stream
    world = dist.get_world_size() if dist is not None else 1
    d_in = torch.from_numpy(big).to(dev)
    d_out = torch.empty(n * 12, dtype=torch.uint8, device=dev)
    box = [None]

    def step(i):
        box[0] = pdist.sharded_voxelgrid(d_in.data_ptr(), n, LEAF, CHUNK, rank, world, d_out.data_ptr(), device=device,
                                         stream=stream)

    step(0)
    step(1)
    ms = timed_region(dist, torch, step, 5)
    n_local, counts, (lo, hi) = box[0]
    report = profile_kernels(pg, torch, step, 2) or {}
    kernels = {k: round(v["total_ms"] / 2, 4) for k, v in report.items()}
    vg = pg.VoxelGrid(LEAF, CHUNK, device=device)
    full = [0]

    def whole(i):
        full[0] = vg.filter_dev(d_in.data_ptr(), n, 12, (0, 4, 8), d_out.data_ptr(), stream)

    whole(0)
    ms1 = timed_region(dist, torch, whole, 5)
    return {"points": n, "world": world, "ms_per_step": ms / 5, "value_mpts": n / (ms / 5 / 1e3) / 1e6,
            "counts_per_rank": counts, "voxels_out": int(sum(counts)), "rank0_chunk_range": [int(lo), int(hi)],
            "unsharded_ms_per_step": ms1 / 5, "unsharded_voxels_out": int(full[0]), "scaling": "strong",
            "rank0_kernel_ms_per_step": kernels}


def bench_icp_sharded(pg, torch, dist, rank, args, peak):
    """BASELINE config 5 (ICP part): ONE alignment of a 1M-pt scan against a 1M-pt base, target sharded over the
    ranks, base index replicated, 16 float64 sums all-reduced over NCCL each iteration."""
    from pcgol_b200 import dist as pdist, synth

    world = args.gpus
    scan = cached_scan(2, N_AZ_1M)
    sensor = np.array([0.0, 0.0, synth.SENSOR_Z], np.float32) + (scan.min(axis=0) * 0)
    target = synth.rigid(scan, 2.0, (0.1, 0.1, 0.05), scan.mean(axis=0))
    lo, hi = pdist.shard_bounds(len(target), rank, world)
    dev = torch.device("cuda")
    device = torch.cuda.current_device()
    stream = torch.cuda.current_stream().cuda_stream
    d_b = torch.from_numpy(scan).to(dev)
    d_t = torch.from_numpy(np.ascontiguousarray(target[lo:hi])).to(dev)
    idx = pg.Index.from_device(d_b.data_ptr(), len(scan), device=device, stream=stream)
    icp = pg.PointToPointICPGradient(pg.PointToPointEvaluator(pg.NearestPointCorresponder(1.0), mode=pg.FAST))
    p = icp.params()
    partial = pdist.make_gpu_partial(idx, d_t.data_ptr(), hi - lo, 1.0, stream)
    res = {}

    def step(i):  # loop resident on the device: partial -> NCCL all-reduce -> finish, no host round trip per iteration
        res["r"] = pdist.sharded_icp_fit_device(idx, d_t.data_ptr(), hi - lo, p, stream=stream)

    def step_host(i):  # the host-driven loop (pcg_icp_partial_dev + pcg_icp_finish), kept for comparison
        res["h"] = pdist.sharded_icp_fit(partial, p)

    step(0)
    step_host(0)
    steps = 3
    ms = timed_region(dist, torch, step, steps)
    ms_host = timed_region(dist, torch, step_host, steps)
    status, trans, stat = res["r"]
    same = bool(np.array_equal(trans, res["h"][1]))
    return {"metric": "sharded ICP alignments/s", "unit": "alignments/s",
            "config": {"workload": "one ICP Fit of a 1M-pt scan (2 deg / 0.15 m perturbed) vs the 1M-pt scan; target "
                                   "split across ranks, index replicated, per-iteration all-reduce of 16 f64 (NCCL), "
                                   "loop resident on the device"},
            "value": steps / (ms / 1e3), "ms_per_alignment": ms / steps, "iterations": int(stat.num_iteration),
            "host_driven_loop": {"value": steps / (ms_host / 1e3), "ms_per_alignment": ms_host / steps,
                                 "same_transform": same},
            "scaling": "strong", "status": int(status), "trans": [float(x) for x in trans]}


def cpu_extras(nn, icp, threads_all):
    """Bounded CPU samples of the other two metrics (oracle as the timed baseline) + parity spot checks."""
    from oracle import oracle as orc

    out = {}
    if nn is not None:
        target, q, ids, dsq = nn.pop("_check")
        t0 = time.perf_counter()
        kdt = orc.Search(target, "kdtree")
        build_s = time.perf_counter() - t0
        sample = 300_000
        t0 = time.perf_counter()
        eids, edsq = kdt.nearest(q[:sample], 1.0, threads=1)
        dt1 = time.perf_counter() - t0
        t0 = time.perf_counter()
        kdt.nearest(q[sample:2 * sample], 1.0, threads=threads_all)
        dtn = time.perf_counter() - t0
        nn["cpu_baseline"] = {"value": sample / dt1, "unit": "queries/s", "cores": 1, "kind": "port",
                              "sample": f"first {sample} of the 10M queries, KD-tree restatement (kdtree.go:83-146), "
                                        f"tree build {build_s:.2f} s not included",
                              "all_cores": {"value": sample / dtn, "cores": threads_all}}
        nn["parity_sample"] = {"queries": sample,
                               "id_mismatches": int((eids != ids[:sample]).sum()),
                               "dist_sq_bit_mismatches": int((edsq.view(np.uint32) != dsq[:sample].view(np.uint32)).sum())}
    if icp is not None:
        base, target = icp.pop("_check")
        kdt = orc.Search(base, "kdtree")
        t0 = time.perf_counter()
        rc, etrans, eev, eit = orc.icp_fit(kdt, target, orc.icp_params(1.0))
        dt = time.perf_counter() - t0
        icp["cpu_baseline"] = {"value": 1.0 / dt, "unit": "alignments/s", "cores": 1, "kind": "port",
                               "sample": f"one full Fit ({eit} iterations), sequential restatement of icp.go:23-67 over "
                                         "the KD-tree restatement"}
        strict = np.array(icp["modes"]["strict"]["trans"], np.float32)
        fast = np.array(icp["modes"]["fast"]["trans"], np.float32)
        icp["parity"] = {"strict_trans_bit_exact": bool(strict.tobytes() == etrans.tobytes()),
                         "strict_iterations_equal": bool(icp["modes"]["strict"]["iterations"] == eit),
                         "fast_max_abs_diff_vs_reference_order": float(np.abs(fast - etrans).max())}
    return out


def run_ours(args):
    import torch

    import pcgol_b200 as pg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    if world != args.gpus:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}; using WORLD_SIZE")
        args.gpus = world
    peak = measured_peak()
    if args.only:
        r = {"nn": bench_nn, "icp": bench_icp, "farm": bench_icp_farm, "vgshard": bench_voxelgrid_sharded}[args.only](pg, torch, dist, rank, args, peak)
        r.pop("_check", None)
        if rank == 0:
            emit({"profiling_aid": args.only, **r})
        return
    vg = bench_voxelgrid(pg, torch, dist, rank, args, peak)
    scan = vg.pop("scan")
    extra = {}
    nn = icp = icp_sh = icp_farm = None
    if not args.no_extra:
        nn = bench_nn(pg, torch, dist, rank, args, peak)
        icp = bench_icp(pg, torch, dist, rank, args, peak)
        icp_sh = bench_icp_sharded(pg, torch, dist, rank, args, peak)
        icp_farm = bench_icp_farm(pg, torch, dist, rank, args, peak)
    cfg5 = bench_config5(pg, torch, dist, rank, args, peak) if args.config5 else None
    line = None
    if rank == 0:
        cores = os.cpu_count() or 1
        # CPU baseline of the primary metric: bounded sample, one core (the reference is single-threaded per call)
        reps = 3
        v1, dt1 = cpu_voxelgrid(scan, 1, reps)
        cpu = {"value": v1, "unit": "Mpts/s", "cores": 1, "kind": "port",
               "sample": f"{reps} full 1M-pt Filter calls ({dt1:.1f} s), C++ restatement of voxelgrid.go:35-187 "
                         "(dense voxel array per chunk), not Go: no Go toolchain in the image"}
        if not args.no_extra:
            cpu_extras(nn, icp, cores)
            extra = {"nn": nn, "icp": icp, "icp_sharded": icp_sh, "icp_farm": icp_farm}
        if cfg5 is not None:
            extra["config5_50m"] = cfg5
        line = {
            "metric": "VoxelGrid Mpts/s", "value": vg["value"], "unit": "Mpts/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": vg["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "points_per_gpu_per_step": 1_000_000, "voxels_out": vg["voxels_out"],
                       "l2": f"input rotates over {ROTATE} distinct device copies (192 MB > 126 MB L2)",
                       "sharding": "one independent cloud per GPU per step, no data-path collective"},
            "e2e": vg["e2e"], "gpu_launches": vg["gpu_launches"], "clocks": vg["clocks"], "roofline": vg["roofline"],
            "pipeline_roofline": vg["pipeline_roofline"], "kernels": vg["kernels"], "cpu_baseline": cpu,
            "host_cores": cores, "extra": extra,
        }
    else:
        if nn is not None:
            nn.pop("_check", None)
        if icp is not None:
            icp.pop("_check", None)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        emit(line)


def main():
    # Libraries (NCCL prints its version) write to fd 1; the contract is ONE JSON line on stdout.
    # Keep the real stdout for that line and point fd 1 at stderr for everything else.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="primary VoxelGrid line only")
    ap.add_argument("--config5", action="store_true", help="add the 50M-point map workload (slow; not in the default run)")
    ap.add_argument("--only", default=None, choices=["nn", "icp", "farm", "vgshard"],
                    help="profiling aid: run just this extra workload and print its object (not a bench line)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
