"""CPU restatement of the reference's PCD (un)marshalling, pc/io.go — TEST INFRASTRUCTURE ONLY
(see oracle/oracle.py).  Pure Python / numpy: the inputs of the tests are small.

Pinned by tests/test_oracle_golden.py against the golden vectors of the reference's own test
(pc/io_test.go:16-216, extracted into tests/golden/pcd_cases.json by tools/gen_golden_pcd.py).

LZF: the reference imports github.com/zhuyie/golzf v0.0.0-20161112031142-8387b0307ade (go.mod:5), a
Go port of liblzf that is NOT vendored in /root/reference.  lzf_decompress below restates the
published liblzf 3.6 format (lzf_d.c): control byte < 32 = literal run of ctrl+1 bytes; otherwise a
back reference of length (ctrl >> 5) + 2 (+ next byte when the 3-bit length is 7) at distance
((ctrl & 0x1f) << 8 | next byte) + 1.  It is anchored on the reference's BinaryCompressed vector.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List

import numpy as np


class PcdSyntaxError(ValueError):  # strconv.ErrSyntax / header validation errors
    pass


class PcdEOF(EOFError):  # io.EOF / io.ErrUnexpectedEOF
    pass


class PcdCorrupt(ValueError):  # lzf.ErrDataCorruption / "wrong uncompressed size"
    pass


class PcdWouldPanic(IndexError):  # the reference indexes a slice out of range
    pass


@dataclass
class Header:  # pc/pointcloud.go:9-18
    version: float = 0.0
    fields: List[str] = field(default_factory=list)
    size: List[int] = field(default_factory=list)
    type: List[str] = field(default_factory=list)
    count: List[int] = field(default_factory=list)
    width: int = 0
    height: int = 0
    viewpoint: List[float] = field(default_factory=list)

    def stride(self) -> int:  # pc/pointcloud.go:64-70
        return sum(s * c for s, c in zip(self.size, self.count))


def lzf_decompress(src: bytes, out_len: int) -> bytes:
    out = bytearray(out_len)
    ip = op = 0
    n = len(src)
    while ip < n:
        ctrl = src[ip]
        ip += 1
        if ctrl < 32:
            ctrl += 1
            if op + ctrl > out_len:
                raise PcdCorrupt("insufficient buffer")
            if ip + ctrl > n:
                raise PcdCorrupt("data corruption")
            out[op:op + ctrl] = src[ip:ip + ctrl]
            ip += ctrl
            op += ctrl
        else:
            length = ctrl >> 5
            ref = op - ((ctrl & 0x1F) << 8) - 1
            if ip >= n:
                raise PcdCorrupt("data corruption")
            if length == 7:
                length += src[ip]
                ip += 1
                if ip >= n:
                    raise PcdCorrupt("data corruption")
            ref -= src[ip]
            ip += 1
            if op + length + 2 > out_len:
                raise PcdCorrupt("insufficient buffer")
            if ref < 0:
                raise PcdCorrupt("data corruption")
            for _ in range(length + 2):
                out[op] = out[ref]
                op += 1
                ref += 1
    return bytes(out[:op])


def _atoi(s: str) -> int:
    try:
        if "_" in s or s.strip() != s:
            raise ValueError
        return int(s, 10)
    except ValueError:
        raise PcdSyntaxError(f'strconv.Atoi: parsing "{s}": invalid syntax')


def _parse_f32(s: str) -> float:
    try:
        if "_" in s or s.strip() != s:
            raise ValueError
        return float(np.float32(float(s)))
    except ValueError:
        raise PcdSyntaxError(f'strconv.ParseFloat: parsing "{s}": invalid syntax')


class _Reader:  # bufio.Reader over a byte string
    def __init__(self, b: bytes):
        self.b, self.pos = b, 0

    def read_line(self) -> bytes:  # ReadLine: strips \n or \r\n; io.EOF when nothing is left
        if self.pos >= len(self.b):
            raise PcdEOF("EOF")
        end = self.b.find(b"\n", self.pos)
        if end < 0:
            line, self.pos = self.b[self.pos:], len(self.b)
        else:
            line, self.pos = self.b[self.pos:end], end + 1
        return line[:-1] if line.endswith(b"\r") else line

    def read_full(self, n: int) -> bytes:
        if self.pos + n > len(self.b):
            raise PcdEOF("unexpected EOF" if self.pos < len(self.b) else "EOF")
        out = self.b[self.pos:self.pos + n]
        self.pos += n
        return out


def unmarshal_header(r: _Reader):
    """unmarshalPCDHeaderTo (io.go:47-135) -> (Header, nPoints, format)."""
    h = Header()
    n_points, fmt = 0, None
    while True:
        args = r.read_line().decode("utf-8", "replace").split()
        if len(args) < 2:
            raise PcdSyntaxError("header field must have value")
        k = args[0]
        if k == "VERSION":
            h.version = _parse_f32(args[1])
        elif k == "FIELDS":
            h.fields = args[1:]
        elif k == "SIZE":
            h.size = [_atoi(s) for s in args[1:]]
        elif k == "TYPE":
            h.type = args[1:]
        elif k == "COUNT":
            h.count = [_atoi(s) for s in args[1:]]
        elif k == "WIDTH":
            h.width = _atoi(args[1])
        elif k == "HEIGHT":
            h.height = _atoi(args[1])
        elif k == "VIEWPOINT":
            h.viewpoint = [_parse_f32(s) for s in args[1:]]
        elif k == "POINTS":
            n_points = _atoi(args[1])
        elif k == "DATA":
            if args[1] not in ("ascii", "binary", "binary_compressed"):
                raise PcdSyntaxError("unknown data format")
            fmt = args[1]
            break
    if len(h.fields) != len(h.size):
        raise PcdSyntaxError("size field size is wrong")
    if len(h.fields) != len(h.type):
        raise PcdSyntaxError("type field size is wrong")
    if len(h.fields) != len(h.count):
        raise PcdSyntaxError("count field size is wrong")
    return h, n_points, fmt


def unmarshal(pcd: bytes):
    """pc.Unmarshal (io.go:32-45,137-230) -> (Header, points, data bytes)."""
    r = _Reader(pcd)
    h, points, fmt = unmarshal_header(r)
    stride = h.stride()
    if fmt == "ascii":
        data = bytearray(points * stride)
        off = 0
        while True:
            try:
                line = r.read_line()
            except PcdEOF:
                break
            tok = line.decode("utf-8", "replace").split()
            lo = 0
            for i, t in enumerate(h.type):
                for j in range(h.count[i]):
                    if t in ("F", "U"):
                        if lo + j >= len(tok):
                            raise PcdWouldPanic("index out of range")
                        if t == "F":
                            v = struct.pack("<f", _parse_f32(tok[lo + j]))
                        else:
                            s = tok[lo + j]
                            if not s.isdigit() or int(s) >= 1 << 32:
                                raise PcdSyntaxError(f'strconv.ParseUint: parsing "{s}": invalid syntax')
                            v = struct.pack("<I", int(s))
                        if off + 4 > len(data):
                            raise PcdWouldPanic("slice bounds out of range")
                        data[off:off + 4] = v
                    off += h.size[i]
                lo += h.count[i]
        return h, points, bytes(data)
    if fmt == "binary":
        return h, points, r.read_full(points * stride)
    n_comp = struct.unpack("<i", r.read_full(4))[0]
    n_unc = struct.unpack("<i", r.read_full(4))[0]
    if n_comp < 0 or n_unc < 0:
        raise PcdWouldPanic("makeslice: len out of range")
    dec = lzf_decompress(r.read_full(n_comp), n_unc)
    if len(dec) != n_unc:
        raise PcdCorrupt("wrong uncompressed size")
    head, offset, pos, off = [], [], 0, 0
    for i in range(len(h.fields)):
        head.append(pos)
        offset.append(off)
        pos += h.size[i] * h.count[i] * points
        off += h.size[i] * h.count[i]
    data = bytearray(n_unc)
    for p in range(points):
        for i in range(len(head)):
            size = h.size[i]
            to, frm = p * stride + offset[i], head[i] + p * size  # sic: `size`, not size*count (io.go:222-224)
            if to + size > n_unc or frm + size > n_unc:
                raise PcdWouldPanic("slice bounds out of range")
            data[to:to + size] = dec[frm:frm + size]
    return h, points, bytes(data)


def marshal(h: Header, points: int, data: bytes) -> bytes:
    """pc.Marshal (io.go:232-285): always DATA binary; an empty Viewpoint gets the default."""
    vp = h.viewpoint if len(h.viewpoint) else [0, 0, 0, 1, 0, 0, 0]
    head = ("VERSION %0.1f\nFIELDS %s\nSIZE %s\nTYPE %s\nCOUNT %s\nWIDTH %d\nHEIGHT %d\nVIEWPOINT %s\nPOINTS %d\n"
            "DATA binary\n") % (float(np.float32(h.version)), " ".join(h.fields), " ".join(str(s) for s in h.size),
                                " ".join(h.type), " ".join(str(c) for c in h.count), h.width, h.height,
                                " ".join("%.4f" % float(np.float32(v)) for v in vp), points)
    return head.encode() + bytes(data)
