"""ctypes front-end of the CPU oracle (oracle/pcgol_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(pcgol_b200/) never imports this module.

Every function mirrors one reference symbol; the C++ side cites file:line.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

OK = 0
E_NO_POINT = 1
E_REF_WOULD_PANIC = 2
E_REF_UNDEFINED = 3
E_TOO_LARGE = 4
E_NOT_ENOUGH_PAIRS = 5


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "pcgol_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_search_new.restype = C.c_void_p
        _lib.orc_search_new.argtypes = [C.c_void_p, C.c_int64, C.c_int32]
        _lib.orc_search_free.argtypes = [C.c_void_p]
        _lib.orc_search_set_min_dist_sq.argtypes = [C.c_void_p, C.c_float]
        _lib.orc_search_delete_point.restype = C.c_int32
        _lib.orc_search_delete_point.argtypes = [C.c_void_p, C.c_int64]
        _lib.orc_search_find_minimum.restype = C.c_int64
        _lib.orc_search_find_minimum.argtypes = [C.c_void_p, C.c_int32]
        _lib.orc_kdtree_num_nodes.restype = C.c_int64
        _lib.orc_kdtree_num_nodes.argtypes = [C.c_void_p]
        _lib.orc_kdtree_max_depth.restype = C.c_int32
        _lib.orc_kdtree_max_depth.argtypes = [C.c_void_p]
        _lib.orc_kdtree_dump.restype = C.c_int32
        _lib.orc_kdtree_dump.argtypes = [C.c_void_p] * 5
        _lib.orc_kdtree_search_leaf.restype = C.c_int64
        _lib.orc_kdtree_search_leaf.argtypes = [C.c_void_p, C.c_void_p]
        _lib.orc_search_nearest.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_int32]
        _lib.orc_search_range.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_voxelgrid_filter.restype = C.c_int32
        _lib.orc_voxelgrid_filter.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]
        _lib.orc_minmax.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_icp_pairs.restype = C.c_int64
        _lib.orc_icp_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_icp_evaluate.restype = C.c_int32
        _lib.orc_icp_evaluate.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_int32, C.c_int32,
                                          C.c_void_p, C.c_void_p]
        _lib.orc_icp_evaluate_w.restype = C.c_int32
        _lib.orc_icp_evaluate_w.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_float, C.c_void_p, C.c_void_p]
        _lib.orc_icp_update.restype = C.c_int32
        _lib.orc_icp_update.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        _lib.orc_icp_fit.restype = C.c_int32
        _lib.orc_icp_fit.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_icp_normal_equations.restype = C.c_int32
        _lib.orc_icp_normal_equations.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_int32, C.c_void_p,
                                                  C.c_void_p, C.c_void_p]
        _lib.orc_icp_fit_gn.restype = C.c_int32
        _lib.orc_icp_fit_gn.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_region_growing_segment.restype = C.c_int64
        _lib.orc_region_growing_segment.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        _lib.orc_mat4_mul.argtypes = [C.c_void_p] * 3
        _lib.orc_mat4_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        _lib.orc_translate.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p]
        _lib.orc_rotate.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]
        _lib.orc_rodrigues.argtypes = [C.c_void_p, C.c_void_p]
        _lib.orc_norm_sq.restype = C.c_float
        _lib.orc_norm_sq.argtypes = [C.c_void_p]
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class IcpParams(C.Structure):
    """Mirrors the knobs of NearestPointCorresponder / PointToPointEvaluator /
    GradientDescentUpdaterFactory; zero means the reference default."""

    _fields_ = [
        ("max_dist", C.c_float),
        ("min_pairs", C.c_int32),
        ("weight", C.c_float * 6),
        ("threshold", C.c_float * 6),
        ("max_iteration", C.c_int32),
        ("f64_accumulate", C.c_int32),
        ("weight_fn", C.c_int32),     # EvaluateWeightFn family: 0 constant (default), 1 truncated, 2 Huber
        ("weight_param", C.c_float),
    ]


def icp_params(max_dist, min_pairs=0, weight=None, threshold=None, max_iteration=0, f64_accumulate=False,
               weight_fn=0, weight_param=0.0) -> IcpParams:
    p = IcpParams()
    p.max_dist = max_dist
    p.min_pairs = min_pairs
    for k in range(6):
        p.weight[k] = 0.0 if weight is None else weight[k]
        p.threshold[k] = 0.0 if threshold is None else threshold[k]
    p.max_iteration = max_iteration
    p.f64_accumulate = 1 if f64_accumulate else 0
    p.weight_fn = weight_fn
    p.weight_param = weight_param
    return p


class Search:
    """storage.Search over a flat xyz array. kind='kdtree' (kdtree.New) or 'naive' (naiveSearch)."""

    def __init__(self, xyz, kind: str = "kdtree", min_dist_sq: float = 0.0):
        self.xyz = _f32(xyz).reshape(-1, 3)
        self.kind = kind
        self._h = lib().orc_search_new(_p(self.xyz), len(self.xyz), 0 if kind == "kdtree" else 1)
        if min_dist_sq:
            lib().orc_search_set_min_dist_sq(self._h, min_dist_sq)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_search_free(self._h)
            self._h = None

    def __len__(self):
        return len(self.xyz)

    def set_min_dist_sq(self, v: float):
        lib().orc_search_set_min_dist_sq(self._h, v)

    def delete_point(self, pid: int) -> bool:
        """KDTree.DeletePoint (kdtree.go:322-332); False == the reference's error (id out of range)."""
        return lib().orc_search_delete_point(self._h, int(pid)) == 0

    def find_minimum(self, dim: int) -> int:
        """findMinimumImpl from the root (kdtree.go:224-266); -2 == the reference's error (dim > 2)."""
        return int(lib().orc_search_find_minimum(self._h, dim))

    def nearest(self, q, max_range: float, threads: int = 1):
        q = _f32(q).reshape(-1, 3)
        ids = np.empty(len(q), np.int64)
        dsq = np.empty(len(q), np.float32)
        lib().orc_search_nearest(self._h, _p(q), len(q), max_range, _p(ids), _p(dsq), threads)
        return ids, dsq

    def range(self, q, max_range: float):
        """CSR (offsets[nq+1], ids, dist_sq); each list in (DistSq, ID) order."""
        q = _f32(q).reshape(-1, 3)
        offsets = np.empty(len(q) + 1, np.int64)
        lib().orc_search_range(self._h, _p(q), len(q), max_range, _p(offsets), None, None)
        total = int(offsets[-1])
        ids = np.empty(total, np.int64)
        dsq = np.empty(total, np.float32)
        lib().orc_search_range(self._h, _p(q), len(q), max_range, _p(offsets), _p(ids), _p(dsq))
        return offsets, ids, dsq

    # KD-tree introspection (golden tree test)
    def dump(self):
        n = lib().orc_kdtree_num_nodes(self._h)
        ids = np.empty(n, np.int64)
        dim = np.empty(n, np.int32)
        left = np.empty(n, np.int32)
        right = np.empty(n, np.int32)
        root = lib().orc_kdtree_dump(self._h, _p(ids), _p(dim), _p(left), _p(right))
        return root, ids, dim, left, right

    def max_depth(self) -> int:
        return lib().orc_kdtree_max_depth(self._h)

    def search_leaf(self, p) -> int:
        p = _f32(p)
        return lib().orc_kdtree_search_leaf(self._h, _p(p))


def voxelgrid_filter(data: np.ndarray, stride: int, off, leaf, chunk=(0, 0, 0), mode: str = "dense",
                     cap_voxels: int = 1 << 28):
    """filter.VoxelGrid on an interleaved record buffer (uint8[n*stride]).

    Returns (status, out_bytes).  mode 'dense' is the literal reference algorithm,
    'sparse' the memory-light equivalent."""
    data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
    n = len(data) // stride if stride else 0
    off = np.asarray(off, np.int64)
    leaf = _f32(leaf)
    chunk = np.asarray(chunk, np.int64)
    out = np.empty(max(1, n * stride), np.uint8)
    n_out = C.c_int64(0)
    rc = lib().orc_voxelgrid_filter(_p(data), n, stride, _p(off), _p(leaf), _p(chunk), 0 if mode == "dense" else 1,
                                    cap_voxels, _p(out), C.byref(n_out))
    return rc, out[: n_out.value * stride].copy()


def minmax(data: np.ndarray, stride: int, off):
    data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
    off = np.asarray(off, np.int64)
    mn = np.empty(3, np.float32)
    mx = np.empty(3, np.float32)
    lib().orc_minmax(_p(data), len(data) // stride, stride, _p(off), _p(mn), _p(mx))
    return mn, mx


def icp_pairs(base: Search, target, max_dist: float):
    t = _f32(target).reshape(-1, 3)
    b = np.empty(len(t), np.int64)
    ti = np.empty(len(t), np.int64)
    d = np.empty(len(t), np.float32)
    n = lib().orc_icp_pairs(base._h, _p(t), len(t), max_dist, _p(b), _p(ti), _p(d))
    return b[:n].copy(), ti[:n].copy(), d[:n].copy()


def icp_evaluate(base: Search, target, max_dist: float, min_pairs: int = 0, f64_accumulate: bool = False,
                 weight_fn: int = 0, weight_param: float = 0.0):
    """Returns (status, ev8={Value, Gradient[6], DistRMS}, n_pairs)."""
    t = _f32(target).reshape(-1, 3)
    out = np.zeros(8, np.float32)
    npairs = C.c_int64(0)
    rc = lib().orc_icp_evaluate_w(base._h, _p(t), len(t), max_dist, min_pairs, 1 if f64_accumulate else 0,
                                  weight_fn, weight_param, _p(out), C.byref(npairs))
    return rc, out, npairs.value


def icp_update(params: IcpParams, i: int, trans16, ev8):
    t = _f32(trans16).copy()
    e = _f32(ev8)
    conv = lib().orc_icp_update(C.byref(params), i, _p(t), _p(e))
    return t, bool(conv)


def icp_fit(base: Search, target, params: IcpParams):
    """Returns (status, trans16 column-major, ev8, num_iteration)."""
    t = _f32(target).reshape(-1, 3)
    trans = np.zeros(16, np.float32)
    ev = np.zeros(8, np.float32)
    it = C.c_int32(0)
    rc = lib().orc_icp_fit(base._h, _p(t), len(t), C.byref(params), _p(trans), _p(ev), C.byref(it))
    return rc, trans, ev, it.value


def icp_normal_equations(base: Search, target, max_dist: float, min_pairs: int = 0):
    """Extension check (not in the reference): (status, hessian36 = 2f*sum J^T J, b6 = sum J^T r, n_pairs)."""
    t = _f32(target).reshape(-1, 3)
    h = np.zeros(36, np.float32)
    b = np.zeros(6, np.float64)
    npairs = C.c_int64(0)
    rc = lib().orc_icp_normal_equations(base._h, _p(t), len(t), max_dist, min_pairs, _p(h), _p(b), C.byref(npairs))
    return rc, h, b, npairs.value


def icp_fit_gn(base: Search, target, params: IcpParams):
    """Extension check: Fit with the Gauss-Newton step. Returns (status, trans16, ev8, num_iteration)."""
    t = _f32(target).reshape(-1, 3)
    trans = np.zeros(16, np.float32)
    ev = np.zeros(8, np.float32)
    it = C.c_int32(0)
    rc = lib().orc_icp_fit_gn(base._h, _p(t), len(t), C.byref(params), _p(trans), _p(ev), C.byref(it))
    return rc, trans, ev, it.value


def region_growing_segment(search: Search, labels, p, max_range: float) -> np.ndarray:
    """RegionGrowing.Segment (regiongrowing.go:23-56): point ids in BFS order."""
    lab = np.ascontiguousarray(labels, np.uint32)
    assert len(lab) == len(search)
    pp = _f32(p)
    out = np.empty(max(1, len(lab)), np.int64)
    n = lib().orc_region_growing_segment(search._h, _p(lab), _p(pp), max_range, _p(out))
    return out[:n].copy()


def mat4_mul(a, b):
    a, b = _f32(a), _f32(b)
    out = np.empty(16, np.float32)
    lib().orc_mat4_mul(_p(a), _p(b), _p(out))
    return out


def mat4_transform(m, xyz):
    m = _f32(m)
    x = _f32(xyz).reshape(-1, 3)
    out = np.empty_like(x)
    lib().orc_mat4_transform(_p(m), _p(x), len(x), _p(out))
    return out


def translate(x, y, z):
    out = np.empty(16, np.float32)
    lib().orc_translate(x, y, z, _p(out))
    return out


def rotate(x, y, z, ang):
    out = np.empty(16, np.float32)
    lib().orc_rotate(x, y, z, ang, _p(out))
    return out


def rodrigues(v):
    v = _f32(v)
    out = np.empty(16, np.float32)
    lib().orc_rodrigues(_p(v), _p(out))
    return out


def norm_sq(v) -> float:
    v = _f32(v)
    return float(lib().orc_norm_sq(_p(v)))
