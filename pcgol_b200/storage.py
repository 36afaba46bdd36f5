"""storage.Search on the GPU (pc/storage/search.go:13-17, replaces pc/storage/kdtree).

`Index` keeps the host accessor for Vec3At/Len/RawIndexAt and owns the device index;
Nearest/Range for one point are batch-of-one calls, NearestBatch/RangeBatch are the
calls the hot path uses.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

from . import _lib
from .pc import as_vec3_buffer


@dataclass
class Neighbor:  # storage.Neighbor
    id: int
    dist_sq: float


def _off(off):
    return (C.c_int64 * 3)(*[int(o) for o in off])


class Index:
    """kdtree.New(ra) replacement: `Index(cloud)` where cloud is a PointCloud or (n,3) float32."""

    def __init__(self, cloud, device: int = 0, min_dist_sq: float = 0.0):
        data, n, stride, off = as_vec3_buffer(cloud)
        self._cloud = cloud
        self._xyz = None
        self.device = device
        #: KDTree.MinDistSq (kdtree.go:19-22): > 0 makes nearest / nearest_batch an approximate search
        self.min_dist_sq = float(min_dist_sq)
        self._h = C.c_void_p()
        self._keep = data
        _lib.check(_lib.lib.pcg_index_build(data.ctypes.data, n, stride, _off(off), device, C.byref(self._h)))

    @classmethod
    def from_device(cls, d_ptr: int, n: int, stride: int = 12, off=(0, 4, 8), device: int = 0, stream: int = 0):
        self = cls.__new__(cls)
        self._cloud = None
        self._xyz = None
        self.device = device
        self.min_dist_sq = 0.0
        self._h = C.c_void_p()
        _lib.check(_lib.lib.pcg_index_build_dev(d_ptr, n, stride, _off(off), device, stream, C.byref(self._h)))
        return self

    def replicate(self, device: int) -> "Index":
        """A bit-identical copy of the built index on another device (NVLink peer copy, no rebuild)."""
        other = Index.__new__(Index)
        other._cloud = self._cloud
        other._xyz = self._xyz
        other.device = device
        other.min_dist_sq = self.min_dist_sq
        other._h = C.c_void_p()
        _lib.check(_lib.lib.pcg_index_replicate(self._h, device, C.byref(other._h)))
        return other

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self, "_shared", None) is None:  # a With() copy does not own the handle
                _lib.lib.pcg_index_free(self._h)
            self._h = None

    __del__ = close

    # -- pc.Vec3RandomAccessor ------------------------------------------------
    def __len__(self) -> int:  # Len
        return int(_lib.lib.pcg_index_len(self._h))

    def vec3_at(self, i: int) -> np.ndarray:  # Vec3At
        if self._xyz is None:
            from .pc import PointCloud
            self._xyz = self._cloud.xyz() if isinstance(self._cloud, PointCloud) else np.asarray(
                self._cloud, np.float32).reshape(-1, 3)
        return self._xyz[i]

    def raw_index_at(self, i: int) -> int:  # RawIndexAt
        return i

    def debug_slots(self) -> np.ndarray:
        """Test hook: the index's point slots as (slots, 4) float32 {x, y, z, original id bits}."""
        m = C.c_int64(0)
        _lib.check(_lib.lib.pcg_debug_index_slots(self._h, None, 0, C.byref(m)))
        out = np.empty((max(m.value, 1), 4), np.float32)
        _lib.check(_lib.lib.pcg_debug_index_slots(self._h, out.ctypes.data, m.value, C.byref(m)))
        return out[: m.value]

    def device_bytes(self) -> int:
        return int(_lib.lib.pcg_index_device_bytes(self._h))

    # -- storage.Search ---------------------------------------------------------
    def nearest(self, p, max_range: float) -> Neighbor:  # Search.Nearest
        ids, dsq = self.nearest_batch(np.asarray(p, np.float32).reshape(1, 3), max_range)
        return Neighbor(int(ids[0]), float(dsq[0]))

    def range(self, p, max_range: float) -> List[Neighbor]:  # Search.Range
        off, ids, dsq = self.range_batch(np.asarray(p, np.float32).reshape(1, 3), max_range)
        return [Neighbor(int(i), float(d)) for i, d in zip(ids, dsq)]

    def nearest_batch(self, queries, max_range: float) -> Tuple[np.ndarray, np.ndarray]:
        """(ids int64[nq], dist_sq float32[nq]); miss = (-1, max_range**2).  Exact unless min_dist_sq > 0."""
        data, n, stride, off = as_vec3_buffer(queries)
        out = np.empty(max(n, 1), dtype=np.dtype([("id", "<i8"), ("dist_sq", "<f4"), ("pad", "<u4")]))
        _lib.check(_lib.lib.pcg_index_nearest_approx(self._h, data.ctypes.data, n, stride, _off(off), max_range,
                                                    self.min_dist_sq, out.ctypes.data))
        return out["id"][:n].copy(), out["dist_sq"][:n].copy()

    def with_min_dist_sq(self, min_dist_sq: float) -> "Index":
        """KDTree.With(opts) (kdtree.go:58-65): a shallow copy sharing the device index."""
        import copy
        other = copy.copy(self)
        other.min_dist_sq = float(min_dist_sq)
        other._shared = self  # keeps the owner alive; only the owner frees the handle
        return other

    # -- KDTree.DeletePoint (kdtree.go:322-332) ---------------------------------
    def delete_point(self, pid: int) -> None:
        self.delete_points([pid])

    def delete_points(self, ids) -> None:
        """Tombstones: the points stop matching any search; len() is unchanged. An id outside
        [0, len-1] raises (status INVALID_ARG, the reference's error text) and deletes nothing."""
        ids = np.ascontiguousarray(ids, np.int64)
        _lib.check(_lib.lib.pcg_index_delete_points(self._h, ids.ctypes.data, len(ids)))

    def nearest_dev(self, d_q: int, nq: int, max_range: float, d_ids: int, d_dist_sq: int, stream: int = 0,
                    stride: int = 12, off=(0, 4, 8)):
        _lib.check(_lib.lib.pcg_index_nearest_dev(self._h, d_q, nq, stride, _off(off), max_range, d_ids, d_dist_sq,
                                                 stream))

    def range_batch(self, queries, max_range: float):
        """CSR (offsets int64[nq+1], ids int64[total], dist_sq float32[total]), lists in (DistSq, ID) order.
        Uses the two-call protocol (count, then fill into a caller-owned buffer)."""
        data, n, stride, off = as_vec3_buffer(queries)
        offs = np.zeros(n + 1, np.int64)
        _lib.check(_lib.lib.pcg_index_range_count(self._h, data.ctypes.data, n, stride, _off(off), max_range,
                                                 offs.ctypes.data))
        total = int(offs[-1])
        nb = np.empty(max(total, 1), dtype=np.dtype([("id", "<i8"), ("dist_sq", "<f4"), ("pad", "<u4")]))
        _lib.check(_lib.lib.pcg_index_range_fill(self._h, data.ctypes.data, n, stride, _off(off), max_range,
                                                offs.ctypes.data, nb.ctypes.data))
        return offs, nb["id"][:total].copy(), nb["dist_sq"][:total].copy()

    def range_count_into(self, queries, max_range: float, offsets_ptr: int) -> None:
        """First call of the two-call protocol into a caller-owned (reusable, ideally pinned) int64[nq + 1] buffer."""
        data, n, stride, off = as_vec3_buffer(queries)
        _lib.check(_lib.lib.pcg_index_range_count(self._h, data.ctypes.data, n, stride, _off(off), max_range, offsets_ptr))

    def range_fill_into(self, queries, max_range: float, offsets_ptr: int, neighbors_ptr: int) -> None:
        """Second call: offsets[nq] storage.Neighbor records (16 bytes each) into a caller-owned buffer."""
        data, n, stride, off = as_vec3_buffer(queries)
        _lib.check(_lib.lib.pcg_index_range_fill(self._h, data.ctypes.data, n, stride, _off(off), max_range, offsets_ptr,
                                                neighbors_ptr))

    def range_batch_owned(self, queries, max_range: float):
        """Same result through the one-call form (library-owned CSR, pcg_index_range)."""
        data, n, stride, off = as_vec3_buffer(queries)
        r = C.c_void_p()
        _lib.check(_lib.lib.pcg_index_range(self._h, data.ctypes.data, n, stride, _off(off), max_range, C.byref(r)))
        try:
            total = int(_lib.lib.pcg_range_total(r))
            offs = np.ctypeslib.as_array(C.cast(_lib.lib.pcg_range_offsets(r), C.POINTER(C.c_int64)),
                                         shape=(n + 1,)).copy()
            if total:
                nb = np.ctypeslib.as_array(C.cast(_lib.lib.pcg_range_neighbors(r), C.POINTER(C.c_uint8)),
                                           shape=(total * 16,)).copy()
                nb = nb.view(np.dtype([("id", "<i8"), ("dist_sq", "<f4"), ("pad", "<u4")]))
                ids, dsq = nb["id"].copy(), nb["dist_sq"].copy()
            else:
                ids, dsq = np.empty(0, np.int64), np.empty(0, np.float32)
        finally:
            _lib.lib.pcg_range_free(r)
        return offs, ids, dsq
