"""filter.Filter / voxelgrid.New on the GPU (pc/filter/filter.go:7-9,
pc/filter/voxelgrid/voxelgrid.go:23-33, option.go:7-18)."""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _lib
from .pc import PointCloud


class NoPointError(ValueError):
    """errors.New("no point") of pc.MinMaxVec3 (pc/minmax.go:10-12)."""


class ReferencePanic(IndexError):
    """The reference would panic with an index out of range on this input."""


class VoxelGrid:
    """voxelgrid.New(leafSize, WithChunkSize(chunk)).  `filter` == Filter.Filter."""

    def __init__(self, leaf_size: Sequence[float], chunk_size: Sequence[int] = (0, 0, 0), device: int = 0):
        self.leaf_size = np.asarray(leaf_size, np.float32).reshape(3)
        self.chunk_size = np.asarray(chunk_size, np.int64).reshape(3)
        self.device = device

    def filter(self, pp: PointCloud) -> PointCloud:
        stride = pp.stride()
        try:
            off = pp.xyz_offsets()
        except KeyError as e:  # pc/pointcloud.go:115 "invalid field name"
            raise ValueError(str(e)) from None
        data = pp.data[: pp.points * stride]
        out = np.empty(max(1, pp.points * stride), np.uint8)
        n_out = C.c_int64(0)
        rc = _lib.lib.pcg_voxelgrid_filter(data.ctypes.data, pp.points, stride, (C.c_int64 * 3)(*off),
                                           self.leaf_size.ctypes.data, self.chunk_size.ctypes.data, self.device,
                                           out.ctypes.data, C.byref(n_out))
        if rc == _lib.E_NO_POINT:
            raise NoPointError("no point")
        if rc == _lib.E_REF_WOULD_PANIC:
            raise ReferencePanic(_lib.last_error())
        _lib.check(rc)
        n = n_out.value
        hdr = pp.header.clone()  # voxelgrid.go:119-128,160-166
        hdr.width, hdr.height = n, 1
        return PointCloud(hdr, out[: n * stride].copy(), n)

    def filter_dev(self, d_data: int, n: int, stride: int, off, d_out: int, stream: int = 0) -> int:
        """Device-resident variant: returns the number of output records."""
        n_out = C.c_int64(0)
        _lib.check(_lib.lib.pcg_voxelgrid_filter_dev(d_data, n, stride, (C.c_int64 * 3)(*off),
                                                    self.leaf_size.ctypes.data, self.chunk_size.ctypes.data,
                                                    self.device, d_out, C.byref(n_out), stream))
        return n_out.value
