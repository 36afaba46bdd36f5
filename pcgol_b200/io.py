"""pc.Unmarshal / pc.Marshal (pc/io.go) and a pc.PointCloud that stays resident in HBM, so that
VoxelGrid -> index build -> ICP run without host round trips."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from .pc import PointCloud, PointCloudHeader
from .storage import Index


class PcdSyntaxError(ValueError):  # strconv.ErrSyntax and the header validation errors of io.go:123-133
    pass


class PcdEOF(EOFError):  # io.EOF / io.ErrUnexpectedEOF
    pass


class PcdCorrupt(ValueError):  # lzf.ErrDataCorruption / "wrong uncompressed size"
    pass


class InvalidField(KeyError):  # errors.New("invalid field name")
    pass


def _check(rc: int):
    if rc == _lib.E_PCD_SYNTAX:
        raise PcdSyntaxError(_lib.last_error())
    if rc == _lib.E_PCD_EOF:
        raise PcdEOF(_lib.last_error())
    if rc == _lib.E_PCD_CORRUPT:
        raise PcdCorrupt(_lib.last_error())
    if rc == _lib.E_INVALID_FIELD:
        raise InvalidField(_lib.last_error())
    _lib.check(rc)


class DeviceCloud:
    """A pc.PointCloud whose Data lives in HBM (handle: pcg_cloud)."""

    def __init__(self, handle: C.c_void_p):
        self._h = handle

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib.pcg_cloud_free(self._h)
            self._h = None

    __del__ = close

    @staticmethod
    def upload(cloud: PointCloud, device: int = 0) -> "DeviceCloud":
        h = cloud.header
        ch = _lib.CloudHeader()
        ch.version = h.version
        ch.n_fields = len(h.fields)
        for i, (f, t) in enumerate(zip(h.fields, h.type)):
            ch.fields[i].value = f.encode()
            ch.type[i].value = t.encode()
            ch.size[i] = h.size[i]
            ch.count[i] = h.count[i]
        ch.width, ch.height = h.width, h.height
        ch.n_viewpoint = len(h.viewpoint)
        for i, v in enumerate(h.viewpoint):
            ch.viewpoint[i] = v
        ch.points = cloud.points
        out = C.c_void_p()
        _check(_lib.lib.pcg_cloud_upload(C.byref(ch), cloud.data.ctypes.data, device, C.byref(out)))
        return DeviceCloud(out)

    def _raw_header(self) -> _lib.CloudHeader:
        ch = _lib.CloudHeader()
        _check(_lib.lib.pcg_cloud_get_header(self._h, C.byref(ch)))
        return ch

    @property
    def points(self) -> int:
        return int(self._raw_header().points)

    def header(self) -> PointCloudHeader:
        ch = self._raw_header()
        n = ch.n_fields
        return PointCloudHeader(version=float(np.float32(ch.version)), fields=[ch.fields[i].value.decode() for i in range(n)],
                                size=[int(ch.size[i]) for i in range(n)], type=[ch.type[i].value.decode() for i in range(n)],
                                count=[int(ch.count[i]) for i in range(n)], width=int(ch.width), height=int(ch.height),
                                viewpoint=[float(ch.viewpoint[i]) for i in range(ch.n_viewpoint)])

    def download(self) -> PointCloud:
        ch = self._raw_header()
        data = np.empty(max(1, ch.data_bytes), np.uint8)
        _check(_lib.lib.pcg_cloud_download(self._h, data.ctypes.data, ch.data_bytes))
        return PointCloud(self.header(), data[: ch.data_bytes], int(ch.points))

    def device_ptr(self) -> int:
        return int(_lib.lib.pcg_cloud_device_ptr(self._h) or 0)

    # -- the resident pipeline ----------------------------------------------------------------
    def voxelgrid(self, leaf, chunk=(0, 0, 0)) -> "DeviceCloud":
        """voxelgrid.New(leaf, WithChunkSize(chunk)).Filter(pp) without leaving the device."""
        lf = (C.c_float * 3)(*[float(x) for x in leaf])
        ck = (C.c_int64 * 3)(*[int(x) for x in chunk])
        out = C.c_void_p()
        from .filter import NoPointError, ReferencePanic
        rc = _lib.lib.pcg_cloud_voxelgrid_filter(self._h, lf, ck, C.byref(out))
        if rc == _lib.E_NO_POINT:
            raise NoPointError("no point")
        if rc == _lib.E_REF_WOULD_PANIC:
            raise ReferencePanic(_lib.last_error())
        _check(rc)
        return DeviceCloud(out)

    def index(self, min_dist_sq: float = 0.0) -> Index:
        """kdtree.New(pp.Vec3Iterator()) on the resident records."""
        idx = Index.__new__(Index)
        idx._cloud = None
        idx._xyz = None
        idx.min_dist_sq = float(min_dist_sq)
        idx._h = C.c_void_p()
        _check(_lib.lib.pcg_cloud_index_build(self._h, C.byref(idx._h)))
        idx.device = int(_lib.lib.pcg_index_device(idx._h))
        return idx

    def icp_fit(self, base: Index, icp) -> tuple:
        """icp.Fit(base, target=self): (trans, Stat)."""
        p = icp.params(base)
        trans = np.zeros(16, np.float32)
        stat = _lib.IcpStat()
        rc = _lib.lib.pcg_cloud_icp_fit(base._h, self._h, C.byref(p), trans.ctypes.data, C.byref(stat))
        return icp._finish(rc, trans, stat)


def unmarshal(pcd: bytes, device: int = 0) -> DeviceCloud:
    """pc.Unmarshal (io.go:32-45) straight into HBM."""
    buf = np.frombuffer(bytes(pcd), np.uint8)
    out = C.c_void_p()
    _check(_lib.lib.pcg_pcd_unmarshal(buf.ctypes.data if len(buf) else None, len(buf), device, C.byref(out)))
    return DeviceCloud(out)


def marshal(cloud: DeviceCloud) -> bytes:
    """pc.Marshal (io.go:232-285): header + records, DATA binary."""
    n = C.c_int64(0)
    _lib.lib.pcg_pcd_marshal(cloud._h, None, 0, C.byref(n))
    buf = np.empty(max(1, n.value), np.uint8)
    _check(_lib.lib.pcg_pcd_marshal(cloud._h, buf.ctypes.data, n.value, C.byref(n)))
    return buf[: n.value].tobytes()
