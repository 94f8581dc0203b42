"""Minimal reader for the reference's PV15 container (videos/test.pv), used ONLY by
tests/golden/make_golden.py and the reference-present oracle test to pin the oracle against the
reference's own segmentation output.  TEST INFRASTRUCTURE ONLY.

Layout restated from Application/src/ProcessedVideo/pv.cpp:842-1051 (header), :286-489 (frame),
:1053-1100 (index) and PVBlob.h:296-338 (ShortHorizontalLine); see SURVEY.md s8c.
LZO1X blocks are decoded by the reference's vendored minilzo compiled into oracle/_ref/.
"""
from __future__ import annotations

import ctypes as C
import os
import struct

import numpy as np

from .seg import LINE_DTYPE, Blobs

_HERE = os.path.dirname(os.path.abspath(__file__))


def _lzo():
    so = os.path.join(_HERE, "_ref", "libminilzo.so")
    if not os.path.exists(so):
        raise FileNotFoundError("oracle/_ref/libminilzo.so missing: run `make -C oracle ref` with /root/reference present")
    L = C.CDLL(so)
    L.lzo1x_decompress_safe.argtypes = [C.c_char_p, C.c_ulong, C.c_char_p, C.POINTER(C.c_ulong), C.c_void_p]
    L.lzo1x_decompress_safe.restype = C.c_int
    return L


class PV15:
    def __init__(self, path):
        self.data = open(path, "rb").read()
        d = self.data
        pos = 0

        def cstr():
            nonlocal pos
            e = d.index(b"\0", pos)
            s = d[pos:e].decode()
            pos = e + 1
            return s

        self.version = cstr()
        assert self.version == "PV15", self.version
        self.encoding = cstr()
        self.width, self.height = struct.unpack_from("<HH", d, pos); pos += 4
        self.offsets = struct.unpack_from("<4H", d, pos); pos += 8
        self.conv_start, self.conv_end = struct.unpack_from("<qq", d, pos); pos += 16
        self.source = cstr()
        self.line_size = d[pos]; pos += 1
        (self.num_frames,) = struct.unpack_from("<I", d, pos); pos += 4
        (self.index_offset,) = struct.unpack_from("<Q", d, pos); pos += 8
        (self.timestamp,) = struct.unpack_from("<Q", d, pos); pos += 8
        self.name = cstr()
        ch = 1 if self.encoding in ("gray", "r3g3b2", "binary") else 3
        self.channels = ch
        self.storage_channels = 0 if self.encoding == "binary" else ch      # required_storage_channels: `binary` frames carry no pixel bytes
        n = self.width * self.height * ch
        self.average = np.frombuffer(d, np.uint8, n, pos).reshape(self.height, self.width, ch).copy()
        if ch == 1:
            self.average = self.average[..., 0]
        pos += n
        self.index = np.frombuffer(d, "<u8", self.num_frames, self.index_offset)
        self._lz = _lzo()

    def frame(self, i) -> Blobs:
        d = self.data
        pos = int(self.index[i])
        compressed = d[pos]; pos += 1
        if compressed:
            csize, usize = struct.unpack_from("<II", d, pos); pos += 8
            out = C.create_string_buffer(usize)
            olen = C.c_ulong(usize)
            r = self._lz.lzo1x_decompress_safe(d[pos:pos + csize], csize, out, C.byref(olen), None)
            assert r == 0 and olen.value == usize, (r, olen.value, usize)
            buf = out.raw
            pos = 0
        else:
            buf = d
        ts, n, src = struct.unpack_from("<QHi", buf, pos); pos += 14
        lines, pixels, lo, po = [], [], [0], [0]
        for _ in range(n):
            start_y, flags, nl = struct.unpack_from("<HBH", buf, pos); pos += 5
            raw = np.frombuffer(buf, "<u2", nl * 2, pos).reshape(nl, 2); pos += nl * 4
            x0 = raw[:, 0]
            x1 = raw[:, 1] & 0x7FFF
            eol = (raw[:, 1] >> 15).astype(np.int64)
            y = start_y + np.concatenate([[0], np.cumsum(eol)[:-1]])
            ln = np.zeros(nl, LINE_DTYPE)
            ln["x0"], ln["x1"], ln["y"] = x0, x1, y
            npx = int((x1.astype(np.int64) - x0 + 1).sum()) * self.storage_channels
            pixels.append(np.frombuffer(buf, np.uint8, npx, pos)); pos += npx
            lines.append(ln)
            lo.append(lo[-1] + nl); po.append(po[-1] + npx)
        L = np.concatenate(lines) if lines else np.zeros(0, LINE_DTYPE)
        P = np.concatenate(pixels) if pixels else np.zeros(0, np.uint8)
        return Blobs(L, P, np.array(lo, np.int64), np.array(po, np.int64))
