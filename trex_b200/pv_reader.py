"""PV15 container reader: the tracker-side input format (the counterpart of pv_writer.py; host-side plumbing).

Layout as pv::Header::read / pv::Frame::read_from consume it (Application/src/ProcessedVideo/pv.cpp:842-1051, :286-489; strings are
NUL-terminated, Application/src/commons/common/file/DataFormat.cpp:509-525; cv::Size is 2 x u16, :530-547):
  "PV15\\0", encoding "\\0", u16 W, H, 4 x u16 crop offsets, i64 conversion range start / end, source "\\0", u8 line size (4),
  u32 frame count, u64 index offset, u64 timestamp, name "\\0", W*H*C average image, u64 mask size (+ mask),
  ... frames ..., index table (frame count x u64 file offsets), metadata "\\0".
  frame: u8 compressed; if 1: u32 compressed size, u32 uncompressed size, LZO1X block (decoded by trex_b200/lzo1x.py);
  payload: u64 timestamp (us), u16 n, i32 source index, then per blob u16 start_y, u8 flags, u16 n_lines,
  n_lines x {u16 x0, u16 x1 | eol << 15} (ShortHorizontalLine, C/processing/PVBlob.h:296-338: y advances after a line with eol),
  the blob's pixel bytes; finally u16 n_predictions (+ predictions, not parsed).
Frames come back in the layout the library's own results use (tb_line records + pixel bytes + offsets), so a stored video can be
fed to the tracker-side stages (tb_seg_rethreshold works on frames, the crop / outline stages on blob lists).
"""
from __future__ import annotations

import struct

import numpy as np

from .lzo1x import decompress

LINE_DTYPE = np.dtype([("x0", "<u2"), ("x1", "<u2"), ("y", "<u2"), ("pad", "<u2")])


class PVFrame:
    def __init__(self, timestamp_us, source_index, lines, pixels, line_off, px_off, flags):
        self.timestamp_us, self.source_index = timestamp_us, source_index
        self.lines, self.pixels, self.line_off, self.px_off, self.flags = lines, pixels, line_off, px_off, flags

    def __len__(self):
        return len(self.line_off) - 1

    def blob(self, k):
        return (self.lines[self.line_off[k]:self.line_off[k + 1]], self.pixels[self.px_off[k]:self.px_off[k + 1]])


class PVReader:
    def __init__(self, path):
        self.data = d = open(path, "rb").read()
        pos = 0

        def cstr():
            nonlocal pos
            e = d.index(b"\0", pos)
            s = d[pos:e].decode(errors="replace")
            pos = e + 1
            return s

        self.version = cstr()
        if self.version != "PV15":
            raise ValueError(f"unsupported PV version {self.version!r} (this reader handles PV15)")
        self.encoding = cstr()
        self.width, self.height = struct.unpack_from("<HH", d, pos); pos += 4
        self.crop_offsets = struct.unpack_from("<4H", d, pos); pos += 8
        self.conversion_range = struct.unpack_from("<qq", d, pos); pos += 16
        self.source = cstr()
        self.line_size = d[pos]; pos += 1
        if self.line_size != 4:
            raise ValueError("PV15 stores lines as 4-byte ShortHorizontalLine records")
        (self.num_frames,) = struct.unpack_from("<I", d, pos); pos += 4
        (self.index_offset,) = struct.unpack_from("<Q", d, pos); pos += 8
        (self.timestamp,) = struct.unpack_from("<Q", d, pos); pos += 8
        self.name = cstr()
        self.channels = 3 if self.encoding == "rgb8" else 1          # channels of the average image
        # bytes per blob pixel in the frames: required_storage_channels(meta_encoding) (C/processing/encoding.h:16-27) -- `binary` stores none
        self.storage_channels = 0 if self.encoding == "binary" else self.channels
        n = self.width * self.height * self.channels
        avg = np.frombuffer(d, np.uint8, n, pos).reshape(self.height, self.width, self.channels).copy(); pos += n
        self.average = avg[..., 0] if self.channels == 1 else avg
        (mask_size,) = struct.unpack_from("<Q", d, pos); pos += 8
        self.mask = np.frombuffer(d, np.uint8, mask_size, pos).copy() if mask_size else None
        self.index = np.frombuffer(d, "<u8", self.num_frames, self.index_offset)
        mpos = self.index_offset + 8 * self.num_frames
        e = d.find(b"\0", mpos)
        self.metadata = d[mpos:e if e >= 0 else len(d)].decode(errors="replace")

    def __len__(self):
        return self.num_frames

    def frame(self, i) -> PVFrame:
        d = self.data
        pos = int(self.index[i])
        if d[pos]:
            csize, usize = struct.unpack_from("<II", d, pos + 1)
            buf = decompress(d[pos + 9:pos + 9 + csize], usize)
            if len(buf) != usize:
                raise ValueError("PV15: LZO block decodes to a different size than recorded")
            pos = 0
        else:
            buf = d; pos += 1
        ts, n, src = struct.unpack_from("<QHi", buf, pos); pos += 14
        lines, pixels, lo, po, flags = [], [], [0], [0], []
        for _ in range(n):
            start_y, fl, nl = struct.unpack_from("<HBH", buf, pos); pos += 5
            raw = np.frombuffer(buf, "<u2", nl * 2, pos).reshape(nl, 2); pos += nl * 4
            ln = np.zeros(nl, LINE_DTYPE)
            ln["x0"] = raw[:, 0]; ln["x1"] = raw[:, 1] & 0x7FFF
            eol = (raw[:, 1] >> 15).astype(np.int64)
            ln["y"] = start_y + np.concatenate([[0], np.cumsum(eol)[:-1]]) if nl else 0
            npx = int((ln["x1"].astype(np.int64) - ln["x0"] + 1).sum()) * self.storage_channels
            pixels.append(np.frombuffer(buf, np.uint8, npx, pos)); pos += npx
            lines.append(ln); flags.append(fl)
            lo.append(lo[-1] + nl); po.append(po[-1] + npx)
        return PVFrame(ts, src, np.concatenate(lines) if lines else np.zeros(0, LINE_DTYPE),
                       np.concatenate(pixels) if pixels else np.zeros(0, np.uint8),
                       np.array(lo, np.int64), np.array(po, np.int64), np.array(flags, np.uint8))
