"""ctypes binding of libpcgol_b200.so (the C ABI of include/pcgol_b200.h).

There is no CPU fallback: if the CUDA library is missing this module raises at
import time, and every compute call raises PcgError when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PCG_LIB") or os.path.join(_HERE, "libpcgol_b200.so")  # PCG_LIB: tuning builds only

OK = 0
E_INVALID_ARG = 1
E_NO_POINT = 2
E_REF_WOULD_PANIC = 3
E_REF_UNDEFINED = 4
E_NOT_ENOUGH_PAIRS = 5
E_CUDA = 6
E_NO_DEVICE = 7
E_TOO_LARGE = 8
E_PCD_SYNTAX = 9
E_PCD_EOF = 10
E_PCD_CORRUPT = 11
E_INVALID_FIELD = 12

ICP_STRICT = 0
ICP_FAST = 1
ICP_WITH_HESSIAN = 0x100
UPDATER_GRADIENT_DESCENT = 0
UPDATER_GAUSS_NEWTON = 1


class PcgError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"pcgol_b200 status {status}: {message}")
        self.status = status
        self.message = message


class Neighbor(C.Structure):
    """storage.Neighbor{ID int; DistSq float32} as Go lays it out (pc/storage/search.go:8-11)."""

    _fields_ = [("id", C.c_int64), ("dist_sq", C.c_float), ("pad_", C.c_uint32)]


WEIGHT_CONSTANT, WEIGHT_TRUNCATED, WEIGHT_HUBER = 0, 1, 2


class IcpParams(C.Structure):
    _fields_ = [
        ("max_dist", C.c_float),
        ("min_pairs", C.c_int32),
        ("weight", C.c_float * 6),
        ("threshold", C.c_float * 6),
        ("max_iteration", C.c_int32),
        ("mode", C.c_int32),
        ("min_dist_sq", C.c_float),  # ABI 2: KDTree.MinDistSq of the base search (kdtree.go:19-22)
        ("updater", C.c_int32),      # ABI 2: UPDATER_GRADIENT_DESCENT | UPDATER_GAUSS_NEWTON
        ("weight_fn", C.c_int32),    # ABI 3: EvaluateWeightFn family (evaluator.go:19-23): WEIGHT_CONSTANT | _TRUNCATED | _HUBER
        ("weight_param", C.c_float),
    ]


class Evaluated(C.Structure):
    """icp.Evaluated (pc/registration/icp/evaluator.go:25-30)."""

    _fields_ = [("value", C.c_float), ("gradient", C.c_float * 6), ("hessian", C.c_float * 36),
                ("dist_rms", C.c_float)]


MAX_FIELDS = 32


class CloudHeader(C.Structure):
    """pcg_cloud_header: pc.PointCloudHeader (pc/pointcloud.go:9-18) + Points + len(Data)."""

    _fields_ = [("version", C.c_float), ("n_fields", C.c_int32), ("fields", (C.c_char * 32) * MAX_FIELDS),
                ("type", (C.c_char * 8) * MAX_FIELDS), ("size", C.c_int64 * MAX_FIELDS),
                ("count", C.c_int64 * MAX_FIELDS), ("width", C.c_int64), ("height", C.c_int64),
                ("n_viewpoint", C.c_int32), ("viewpoint", C.c_float * 16), ("points", C.c_int64),
                ("data_bytes", C.c_int64)]


class IcpStat(C.Structure):
    """icp.Stat (pc/registration/icp/stat.go:3-6) + pair count of the last Evaluate."""

    _fields_ = [("evaluated", Evaluated), ("num_iteration", C.c_int32), ("pad_", C.c_int32), ("n_pairs", C.c_int64)]


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a). pcgol_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_vp, _i64, _i32, _f = C.c_void_p, C.c_int64, C.c_int32, C.c_float
_sigs = {
    "pcg_abi_version": (_i32, []),
    "pcg_last_error": (C.c_char_p, []),
    "pcg_status_string": (C.c_char_p, [_i32]),
    "pcg_device_count": (_i32, []),
    "pcg_host_alloc": (_i32, [C.POINTER(_vp), _i64]),
    "pcg_host_free": (None, [_vp]),
    "pcg_device_alloc": (_i32, [_i32, C.POINTER(_vp), _i64]),
    "pcg_device_free": (None, [_i32, _vp]),
    "pcg_memcpy_h2d": (_i32, [_i32, _vp, _vp, _i64]),
    "pcg_memcpy_d2h": (_i32, [_i32, _vp, _vp, _i64]),
    "pcg_device_synchronize": (_i32, [_i32]),
    "pcg_kernel_launch_count": (_i64, []),
    "pcg_debug_sequential_sum_f32": (_i32, [_vp, _i64, _i32, _i32, _vp]),
    "pcg_debug_index_slots": (_i32, [_vp, _vp, _i64, C.POINTER(_i64)]),
    "pcg_profile_enable": (None, [_i32]),
    "pcg_debug_set_vg_path": (None, [_i32]),
    "pcg_profile_report": (_i64, [C.c_char_p, _i64]),
    "pcg_index_build": (_i32, [_vp, _i64, _i64, _vp, _i32, C.POINTER(_vp)]),
    "pcg_index_build_dev": (_i32, [_vp, _i64, _i64, _vp, _i32, _vp, C.POINTER(_vp)]),
    "pcg_index_free": (None, [_vp]),
    "pcg_index_len": (_i64, [_vp]),
    "pcg_index_device": (_i32, [_vp]),
    "pcg_index_device_bytes": (_i64, [_vp]),
    "pcg_index_nearest": (_i32, [_vp, _vp, _i64, _i64, _vp, _f, _vp]),
    "pcg_index_nearest_dev": (_i32, [_vp, _vp, _i64, _i64, _vp, _f, _vp, _vp, _vp]),
    "pcg_index_nearest_approx": (_i32, [_vp, _vp, _i64, _i64, _vp, _f, _f, _vp]),
    "pcg_index_nearest_approx_dev": (_i32, [_vp, _vp, _i64, _i64, _vp, _f, _f, _vp, _vp, _vp]),
    "pcg_index_delete_points": (_i32, [_vp, _vp, _i64]),
    "pcg_index_range": (_i32, [_vp, _vp, _i64, _i64, _vp, _f, C.POINTER(_vp)]),
    "pcg_index_range_count": (_i32, [_vp, _vp, _i64, _i64, _vp, _f, _vp]),
    "pcg_index_range_fill": (_i32, [_vp, _vp, _i64, _i64, _vp, _f, _vp, _vp]),
    "pcg_range_total": (_i64, [_vp]),
    "pcg_range_offsets": (_vp, [_vp]),
    "pcg_range_neighbors": (_vp, [_vp]),
    "pcg_range_free": (None, [_vp]),
    "pcg_pcd_unmarshal": (_i32, [_vp, _i64, _i32, C.POINTER(_vp)]),
    "pcg_pcd_marshal": (_i32, [_vp, _vp, _i64, C.POINTER(_i64)]),
    "pcg_cloud_upload": (_i32, [C.POINTER(CloudHeader), _vp, _i32, C.POINTER(_vp)]),
    "pcg_cloud_get_header": (_i32, [_vp, C.POINTER(CloudHeader)]),
    "pcg_cloud_download": (_i32, [_vp, _vp, _i64]),
    "pcg_cloud_device_ptr": (_vp, [_vp]),
    "pcg_cloud_free": (None, [_vp]),
    "pcg_cloud_voxelgrid_filter": (_i32, [_vp, _vp, _vp, C.POINTER(_vp)]),
    "pcg_cloud_index_build": (_i32, [_vp, C.POINTER(_vp)]),
    "pcg_cloud_icp_fit": (_i32, [_vp, _vp, C.POINTER(IcpParams), _vp, C.POINTER(IcpStat)]),
    "pcg_region_growing_new": (_i32, [_vp, _vp, _i64, _i64, _vp, _i64, C.POINTER(_vp)]),
    "pcg_region_growing_free": (None, [_vp]),
    "pcg_region_growing_segment": (_i32, [_vp, _vp, _f, _vp, _i64, C.POINTER(_i64)]),
    "pcg_voxelgrid_filter": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _i32, _vp, C.POINTER(_i64)]),
    "pcg_voxelgrid_filter_dev": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _i32, _vp, C.POINTER(_i64), _vp]),
    "pcg_voxelgrid_chunk_histogram_dev": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _i32, _i64, _vp, _i64, C.POINTER(_i64),
                                          _vp]),
    "pcg_voxelgrid_filter_chunks_dev": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _i64, _i64, _i32, _vp, C.POINTER(_i64),
                                               _vp]),
    "pcg_minmax_dev": (_i32, [_vp, _i64, _i64, _vp, _i32, _vp, _vp, _vp]),
    "pcg_icp_pairs": (_i32, [_vp, _vp, _i64, _i64, _vp, _f, _vp, _vp, _vp, C.POINTER(_i64)]),
    "pcg_icp_pairs_approx": (_i32, [_vp, _vp, _i64, _i64, _vp, _f, _f, _vp, _vp, _vp, C.POINTER(_i64)]),
    "pcg_icp_evaluate_params": (_i32, [_vp, _vp, _i64, _i64, _vp, C.POINTER(IcpParams), C.POINTER(Evaluated),
                                       C.POINTER(_i64)]),
    "pcg_icp_evaluate": (_i32, [_vp, _vp, _i64, _i64, _vp, _f, _i32, _i32, C.POINTER(Evaluated), C.POINTER(_i64)]),
    "pcg_icp_fit": (_i32, [_vp, _vp, _i64, _i64, _vp, C.POINTER(IcpParams), _vp, C.POINTER(IcpStat)]),
    "pcg_icp_fit_dev": (_i32, [_vp, _vp, _i64, _i64, _vp, C.POINTER(IcpParams), _vp, C.POINTER(IcpStat), _vp]),
    "pcg_icp_fit_pairs_dev": (_i32, [_i32, _vp, _vp, _vp, _vp, _i64, _vp, C.POINTER(IcpParams), _i32, _vp, _vp, _vp,
                                     _vp]),
    "pcg_icp_partial_dev": (_i32, [_vp, _vp, _i64, _i64, _vp, _f, _vp, _i32, _vp, _vp, _vp]),
    "pcg_minmax_packed_dev": (_i32, [_vp, _i64, _i64, _vp, _i32, _i64, _vp, _vp]),
    "pcg_voxelgrid_chunk_histogram_mm_dev": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _i32, _i64, _vp, _i64, C.POINTER(_i64), _vp]),
    "pcg_voxelgrid_owner_order_dev": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "pcg_voxelgrid_filter_chunks_mm_dev": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _vp, C.POINTER(_i64), _vp]),
    "pcg_index_replicate": (_i32, [_vp, _i32, C.POINTER(_vp)]),
    "pcg_icp_fit_multi": (_i32, [_i32, _vp, _vp, _i64, _i64, _vp, C.POINTER(IcpParams), _vp, C.POINTER(IcpStat)]),
    "pcg_icp_fit_multi_dev": (_i32, [_i32, _vp, _vp, _vp, _i64, _vp, C.POINTER(IcpParams), _vp, C.POINTER(IcpStat)]),
    "pcg_icp_shard_new": (_i32, [_vp, _vp, _i64, _i64, _vp, C.POINTER(IcpParams), _vp, C.POINTER(_vp)]),
    "pcg_icp_shard_free": (None, [_vp]),
    "pcg_icp_shard_partial": (_i32, [_vp, _vp, _vp]),
    "pcg_icp_shard_finish": (_i32, [_vp, _vp, _vp]),
    "pcg_icp_shard_result": (_i32, [_vp, _vp, C.POINTER(IcpStat), C.POINTER(_i32), _vp]),
    "pcg_query_order_dev": (_i32, [_vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "pcg_icp_finish": (_i32, [_vp, C.POINTER(IcpParams), C.POINTER(_i32), _vp, C.POINTER(Evaluated),
                              C.POINTER(_i32)]),
}
for _name, (_res, _args) in _sigs.items():
    _fn = getattr(lib, _name)  # AttributeError here == the library does not export a declared symbol
    _fn.restype = _res
    _fn.argtypes = _args

EXPORTED_SYMBOLS = tuple(_sigs)


def last_error() -> str:
    return (lib.pcg_last_error() or b"").decode("utf-8", "replace")


def check(status: int, ok=(OK,)) -> int:
    if status not in ok:
        raise PcgError(status, last_error() or (lib.pcg_status_string(status) or b"").decode())
    return status


def device_count() -> int:
    return int(lib.pcg_device_count())


def kernel_launch_count() -> int:
    return int(lib.pcg_kernel_launch_count())


def profile_enable(on: bool) -> None:
    lib.pcg_profile_enable(1 if on else 0)


def set_vg_path(path: int) -> None:
    """Test hook: 0 automatic, 1 always the multi-kernel pipeline (packed words when they fit), 2 always the
    (key, index) pairs pipeline."""
    lib.pcg_debug_set_vg_path(int(path))


def profile_report() -> dict:
    """Per-kernel {"launches", "total_ms"} recorded since profile_enable(True)."""
    import json
    buf = C.create_string_buffer(1 << 16)
    lib.pcg_profile_report(buf, len(buf))
    return json.loads(buf.value.decode() or "{}")
