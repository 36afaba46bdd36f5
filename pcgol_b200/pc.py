"""Boundary data types: the part of package pc the hot path touches.

PointCloud mirrors pc.PointCloud / pc.PointCloudHeader (pc/pointcloud.go:9-78): an
interleaved little-endian record buffer described by Fields/Size/Type/Count.
A pc.Vec3Slice (pc/vec3slice.go:8) is simply a float32 array of shape (n, 3).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

import numpy as np


@dataclass
class PointCloudHeader:
    version: float = 0.7
    fields: List[str] = field(default_factory=lambda: ["x", "y", "z"])
    size: List[int] = field(default_factory=lambda: [4, 4, 4])
    type: List[str] = field(default_factory=lambda: ["F", "F", "F"])
    count: List[int] = field(default_factory=lambda: [1, 1, 1])
    width: int = 0
    height: int = 1
    viewpoint: List[float] = field(default_factory=lambda: [0, 0, 0, 1, 0, 0, 0])

    def clone(self) -> "PointCloudHeader":  # pc/pointcloud.go:20-31
        return PointCloudHeader(self.version, list(self.fields), list(self.size), list(self.type), list(self.count),
                                self.width, self.height, list(self.viewpoint))

    def stride(self) -> int:  # pc/pointcloud.go:64-70
        return sum(c * s for c, s in zip(self.count, self.size))


class PointCloud:
    def __init__(self, header: PointCloudHeader, data, points: int | None = None):
        self.header = header
        self.data = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        stride = header.stride()
        self.points = len(self.data) // stride if points is None else points
        if self.points * stride > len(self.data):
            raise ValueError("data shorter than points * stride")

    @staticmethod
    def from_xyz(xyz) -> "PointCloud":
        a = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
        h = PointCloudHeader(width=len(a), height=1)
        return PointCloud(h, a.view(np.uint8).reshape(-1), len(a))

    def stride(self) -> int:
        return self.header.stride()

    def field_offset(self, name: str) -> int:  # pc/pointcloud.go:88-116
        off = 0
        for i, fn in enumerate(self.header.fields):
            if fn == name:
                return off
            off += self.header.size[i] * self.header.count[i]
        raise KeyError("invalid field name")  # errors.New("invalid field name")

    def xyz_offsets(self) -> Tuple[int, int, int]:
        """Byte offsets of x, y, z inside a record, resolved like PointCloud.Vec3Iterator
        (pc/pointcloud.go:130-171): "xyz" if it comes before a complete x,y,z run, else consecutive
        x,y,z, else (naiveVec3Iterator) the three fields wherever they are."""
        h = self.header
        state, first = 0, None
        for name in h.fields:
            if name == "xyz":
                state, first = 3, name
                break
            if name == "x" and state == 0:
                state, first = 1, name
            elif name == "y" and state == 1:
                state = 2
            elif name == "z" and state == 2:
                state = 3
                break
            else:
                state = 0
        if state == 3:
            o = self.field_offset(first)
            if self.stride() % 4 == 0 and o % 4 == 0:
                return (o, o + 4, o + 8)
        return (self.field_offset("x"), self.field_offset("y"), self.field_offset("z"))

    def xyz(self) -> np.ndarray:
        """Vec3At for every record -> float32 (n, 3) copy."""
        s = self.stride()
        rec = self.data[: self.points * s].reshape(self.points, s)
        ox, oy, oz = self.xyz_offsets()
        out = np.empty((self.points, 3), np.float32)
        for k, o in enumerate((ox, oy, oz)):
            out[:, k] = rec[:, o:o + 4].copy().view("<f4").reshape(-1)
        return out

    def field_u32(self, name: str) -> np.ndarray:
        s = self.stride()
        o = self.field_offset(name)
        rec = self.data[: self.points * s].reshape(self.points, s)
        return rec[:, o:o + 4].copy().view("<u4").reshape(-1)


def as_vec3_buffer(cloud) -> Tuple[np.ndarray, int, int, Sequence[int]]:
    """Flattens a PointCloud or an (n,3) float32 array (Vec3Slice) to (bytes, n, stride, xyz_off)."""
    if isinstance(cloud, PointCloud):
        return cloud.data, cloud.points, cloud.stride(), cloud.xyz_offsets()
    a = np.ascontiguousarray(cloud, dtype=np.float32).reshape(-1, 3)
    return a.view(np.uint8).reshape(-1), len(a), 12, (0, 4, 8)
