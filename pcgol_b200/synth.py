"""Deterministic synthetic LiDAR-shaped clouds for the parity tests and bench.py
(SURVEY.md §8d).  numpy only; seeds fixed per configuration.

lidar_scan: a 64-beam spinning sensor at (0, 0, 1.73 m) inside an 80 x 50 m walled yard
with a ground plane and 40 random boxes; elevation -24.8 .. +2.0 deg, `n_az` azimuth
steps (1875 -> 120k rays, 15625 -> 1M rays), max range 100 m, range noise N(0, 0.02 m).
Points are emitted azimuth-major (a packet of 64 beams per azimuth step) and the cloud is
translated so that min(x, y, z) == 0 exactly (needed by the reference's un-chunked
VoxelGrid path, pc/filter/voxelgrid/voxelgrid.go:46).
"""
from __future__ import annotations

import numpy as np

SENSOR_Z = 1.73


def _scene(seed: int):
    rng = np.random.default_rng(1000 + seed)
    nb = 40
    centre = np.stack([rng.uniform(-38, 38, nb), rng.uniform(-23, 23, nb)], 1)
    # keep the sensor outside every box
    near = np.linalg.norm(centre, axis=1) < 4.0
    centre[near] += 8.0
    half = np.stack([rng.uniform(0.25, 2.0, nb), rng.uniform(0.25, 2.0, nb)], 1)
    height = rng.uniform(0.5, 3.0, nb)
    lo = np.concatenate([centre - half, np.zeros((nb, 1))], 1)
    hi = np.concatenate([centre + half, height[:, None]], 1)
    return lo, hi


def lidar_scan(seed: int = 0, n_az: int = 1875, beams: int = 64, pose=None, shift_to_zero: bool = True,
               return_sensor: bool = False):
    """float32 (n, 3).  pose = (x, y, yaw) of the sensor in the scene (default origin)."""
    lo, hi = _scene(seed)
    rng = np.random.default_rng(seed)
    px, py, yaw = (0.0, 0.0, 0.0) if pose is None else pose
    elev = np.deg2rad(np.linspace(-24.8, 2.0, beams))
    az = yaw + np.arange(n_az) * (2 * np.pi / n_az)
    A, E = np.meshgrid(az, elev, indexing="ij")  # azimuth-major
    A = A.reshape(-1)
    E = E.reshape(-1)
    d = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], 1)
    o = np.array([px, py, SENSOR_Z])
    t = np.full(len(d), np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = np.where(d[:, 2] < 0, -o[2] / d[:, 2], np.inf)  # ground z = 0
        t = np.minimum(t, tg)
        for axis, (a, b) in enumerate(((-40.0, 40.0), (-25.0, 25.0))):  # yard walls
            tw = np.where(d[:, axis] > 0, (b - o[axis]) / d[:, axis],
                          np.where(d[:, axis] < 0, (a - o[axis]) / d[:, axis], np.inf))
            t = np.minimum(t, np.where(tw > 0, tw, np.inf))
        inv = 1.0 / d
        for k in range(len(lo)):  # boxes: slab test
            t0 = (lo[k] - o) * inv
            t1 = (hi[k] - o) * inv
            tn = np.minimum(t0, t1).max(axis=1)
            tf = np.maximum(t0, t1).min(axis=1)
            hit = (tn <= tf) & (tf > 0) & (tn > 0)
            t = np.where(hit, np.minimum(t, tn), t)
    keep = np.isfinite(t) & (t <= 100.0)
    t = t + rng.normal(0.0, 0.02, len(t))
    pts = o + d * t[:, None]
    pts = pts[keep]
    # sensor frame -> the cloud is expressed relative to the sensor pose
    c, s = np.cos(-yaw), np.sin(-yaw)
    rel = pts - np.array([px, py, 0.0])
    pts = np.stack([c * rel[:, 0] - s * rel[:, 1], s * rel[:, 0] + c * rel[:, 1], rel[:, 2]], 1)
    pts = pts.astype(np.float32)
    sensor = np.array([0.0, 0.0, SENSOR_Z], np.float32)
    if shift_to_zero:
        mn = pts.min(axis=0)
        pts = (pts - mn).astype(np.float32)
        sensor = (sensor - mn).astype(np.float32)
    return (pts, sensor) if return_sensor else pts


def rigid(points: np.ndarray, yaw_deg: float, trans, centre) -> np.ndarray:
    """Rotate about the z axis through `centre`, then translate (float64 math, float32 result)."""
    a = np.deg2rad(yaw_deg)
    c, s = np.cos(a), np.sin(a)
    p = points.astype(np.float64) - np.asarray(centre, np.float64)
    out = np.stack([c * p[:, 0] - s * p[:, 1], s * p[:, 0] + c * p[:, 1], p[:, 2]], 1)
    return (out + np.asarray(centre, np.float64) + np.asarray(trans, np.float64)).astype(np.float32)


def icp_pair(seed: int = 1, n: int = 100_000, n_az: int = 1875, yaw_deg: float = 5.0, trans=(0.2, 0.2, 0.1)):
    """BASELINE config 1: base = n-point subsample of a scan, target = a 5 deg / 0.3 m perturbed copy."""
    scan, sensor = lidar_scan(seed, n_az=n_az, return_sensor=True)
    rng = np.random.default_rng(77 + seed)
    if n < len(scan):
        idx = np.sort(rng.choice(len(scan), n, replace=False))
        scan = scan[idx]
    target = rigid(scan, yaw_deg, trans, sensor)
    return np.ascontiguousarray(scan), np.ascontiguousarray(target)


def scan_pair(k: int, n_az: int = 1875):
    """BASELINE config 4: scan k and the same scene seen from a pose perturbed by <= 5 deg / 0.3 m."""
    rng = np.random.default_rng(1_000_000 + k)
    yaw = np.deg2rad(rng.uniform(-5, 5))
    v = rng.normal(size=2)
    v = v / np.linalg.norm(v) * rng.uniform(0, 0.3)
    base = lidar_scan(k, n_az=n_az, shift_to_zero=False)
    target = lidar_scan(k, n_az=n_az, pose=(v[0], v[1], yaw), shift_to_zero=False)
    return base, target


def nn_queries(target: np.ndarray, nq: int, seed: int = 3, sigma: float = 0.3) -> np.ndarray:
    """BASELINE config 3: queries = target points (cycled, in scan order) + N(0, sigma) jitter."""
    rng = np.random.default_rng(seed)
    reps = -(-nq // len(target))
    q = np.tile(target, (reps, 1))[:nq].astype(np.float32)
    q += rng.normal(0.0, sigma, q.shape).astype(np.float32)
    return np.ascontiguousarray(q)


def with_fields(xyz: np.ndarray, extra_u32: int = 1, seed: int = 9) -> tuple:
    """Interleaves x,y,z with `extra_u32` uint32 fields (label, ...): returns (bytes, stride, xyz_off)."""
    n = len(xyz)
    rng = np.random.default_rng(seed)
    rec = np.empty((n, 3 + extra_u32), np.uint32)
    rec[:, :3] = xyz.view(np.uint32)
    rec[:, 3:] = rng.integers(0, 2**32 - 1, size=(n, extra_u32), dtype=np.uint32)
    return rec.view(np.uint8).reshape(-1), 4 * (3 + extra_u32), (0, 4, 8)


def tiled_map(n_tiles_x: int, n_tiles_y: int, seed: int = 2, n_az: int = 15625, spacing=(78.0, 48.0)) -> np.ndarray:
    """BASELINE config 5 stand-in: a large map made of shifted copies of 1M-point scans (ray casting
    50M rays on the host would take minutes).  Copies get a small per-tile jitter so that no two points
    coincide; min(x, y, z) stays 0."""
    scan = lidar_scan(seed, n_az=n_az)
    rng = np.random.default_rng(seed + 12345)
    out = np.empty((n_tiles_x * n_tiles_y * len(scan), 3), np.float32)
    k = 0
    for ix in range(n_tiles_x):
        for iy in range(n_tiles_y):
            sh = np.array([ix * spacing[0], iy * spacing[1], 0.0], np.float32)
            jit = rng.normal(0.0, 0.003, scan.shape).astype(np.float32) if (ix or iy) else 0.0
            out[k:k + len(scan)] = scan + sh + jit
            k += len(scan)
    out -= out.min(axis=0)
    return out
