"""pc/segmentation/regiongrowing on the GPU (regiongrowing.go:11-56): every breadth-first level is one
batched Range on the device index."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .pc import PointCloud
from .storage import Index


class RegionGrowing:
    """regiongrowing.New(search, propertyIter): `cloud` is the PointCloud the index was built from and
    `field` the uint32 property (pc.Uint32Iterator(field))."""

    def __init__(self, search: Index, cloud: PointCloud, field: str = "label"):
        self._search = search
        off = (C.c_int64 * 3)(*cloud.xyz_offsets())
        self._h = C.c_void_p()
        self._n = cloud.points
        _lib.check(_lib.lib.pcg_region_growing_new(search._h, cloud.data.ctypes.data, cloud.points, cloud.stride(), off,
                                                  cloud.field_offset(field), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib.pcg_region_growing_free(self._h)
            self._h = None

    __del__ = close

    def segment(self, p, max_range: float) -> np.ndarray:
        """Segment(p, maxRange): int64 point ids in the reference's breadth-first order."""
        pp = np.ascontiguousarray(p, np.float32).reshape(3)
        out = np.empty(max(1, self._n), np.int64)
        m = C.c_int64(0)
        _lib.check(_lib.lib.pcg_region_growing_segment(self._h, pp.ctypes.data, max_range, out.ctypes.data, self._n,
                                                      C.byref(m)))
        return out[: m.value].copy()
