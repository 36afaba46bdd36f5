"""Diagnostic: per-rank timings of e2e VoxelGrid (1 and 3 callers) and ICP fast under torchrun."""
import ctypes as C, os, sys, threading, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcgol_b200 as pg
from pcgol_b200 import synth
rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
use_nccl = world > 1 and os.environ.get("NO_NCCL") is None
if use_nccl:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
print(rank, "affinity", len(os.sched_getaffinity(0)), "cpus", flush=True)
scan = synth.lidar_scan(2 + rank, n_az=15625)
n = len(scan)
leaf = np.asarray((0.05,) * 3, np.float32); chunk = np.asarray((128,) * 3, np.int64); off = (C.c_int64 * 3)(0, 4, 8)
def make():
    h_in = torch.from_numpy(scan.view(np.uint8).reshape(-1).copy()).pin_memory()
    h_out = torch.empty(n * 12, dtype=torch.uint8).pin_memory(); n_out = C.c_int64(0)
    def call():
        rc = pg._lib.lib.pcg_voxelgrid_filter(h_in.data_ptr(), n, 12, off, leaf.ctypes.data, chunk.ctypes.data, local, h_out.data_ptr(), C.byref(n_out))
        assert rc == 0, pg._lib.last_error()
    return call
for callers in (1, 3):
    calls = [make() for _ in range(callers)]
    for c in calls:
        for _ in range(3): c()
    times = []
    def worker(c, k):
        for _ in range(10):
            t0 = time.perf_counter(); c(); times.append((k, time.perf_counter() - t0))
    t0 = time.perf_counter()
    th = [threading.Thread(target=worker, args=(c, k)) for k, c in enumerate(calls)]
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
    per = sorted(x[1] for x in times)
    print(f"rank {rank} callers {callers}: total {dt*1e3:.2f} ms for {10*callers} calls; per-call min {per[0]*1e3:.3f} med {per[len(per)//2]*1e3:.3f} max {per[-1]*1e3:.3f}", flush=True)
base, target = synth.icp_pair(seed=1)
d_b = torch.from_numpy(base).cuda(); d_t = torch.from_numpy(target).cuda()
stream = torch.cuda.current_stream().cuda_stream
idx = pg.Index.from_device(d_b.data_ptr(), len(base), device=local, stream=stream)
icp = pg.PointToPointICPGradient(pg.PointToPointEvaluator(pg.NearestPointCorresponder(1.0), mode=pg.FAST))
for rep in range(3):
    icp.fit_dev(idx, d_t.data_ptr(), len(target), stream)
if use_nccl: dist.barrier()
ts = []
for rep in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record(); icp.fit_dev(idx, d_t.data_ptr(), len(target), stream); b.record(); torch.cuda.synchronize()
    ts.append((a.elapsed_time(b), (time.perf_counter() - t0) * 1e3))
print(f"rank {rank} icp fast: event ms {[round(x[0],2) for x in ts]} wall {[round(x[1],2) for x in ts]}", flush=True)
import subprocess
print(rank, subprocess.run(["nvidia-smi", "--query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader", "-i", str(local)], capture_output=True, text=True).stdout.strip(), flush=True)
if use_nccl:
    dist.barrier(); dist.destroy_process_group()
