"""PV15 container writer ("next" row N2): serialises the blob lists the GPU path produces into the
reference's `.pv` format so TRex (`-task track`, pvinfo, ...) can open them.  Host-side plumbing; format restated from
  Header::write        Application/src/ProcessedVideo/pv.cpp:1053-1165 (layout comment :1060-1100)
  Frame::serialize     pv.cpp:666-775 (uncompressed form; TRex itself stores small frames uncompressed, :713)
  ShortHorizontalLine  Application/src/commons/common/processing/PVBlob.h:296-338, PVBlob.cpp:293-316
Frames are written uncompressed (compression flag 0) by default, which every reader of the format accepts
(Frame::read_from, pv.cpp:315-330).  With compress=True a frame payload goes through an LZO1X encoder (trex_b200/lzo1x.py) and is
stored as {u8 1, u32 compressed size, u32 uncompressed size, block} when that is smaller -- the reference's rule
(`size < in_len`, pv.cpp:741-763); the encoder is not minilzo's, so the block bytes differ from a TRex-written file while any
LZO1X decoder reads them.
"""
from __future__ import annotations

import json
import struct
import time

import numpy as np


class PVWriter:
    def __init__(self, path, width, height, average: np.ndarray, *, encoding="gray", source="", name="",
                 conversion_range=(-1, -1), crop_offsets=(0, 0, 0, 0), metadata: dict | None = None, compress=False):
        assert encoding == "gray", "only meta_encoding=gray is built"
        average = np.ascontiguousarray(average, np.uint8)
        assert average.shape == (height, width)
        self.f = open(path, "wb")
        self.width, self.height = width, height
        self.index = []
        self.metadata = dict(metadata or {})
        self.compress = bool(compress)
        self.compressed_frames = 0
        w = self.f.write
        w(b"PV15\0")
        w(encoding.encode() + b"\0")
        w(struct.pack("<HH", width, height))                       # cv::Size as 2 x u16 (DataFormat.cpp:530-547)
        w(struct.pack("<4H", *crop_offsets))
        w(struct.pack("<qq", *conversion_range))
        w(source.encode() + b"\0")
        w(struct.pack("<B", 4))                                     # sizeof(ShortHorizontalLine)
        self._num_frames_offset = self.f.tell()
        w(struct.pack("<I", 0))
        self._index_offset_offset = self.f.tell()
        w(struct.pack("<Q", 0))
        w(struct.pack("<Q", int(time.time() * 1e6)))
        w(name.encode() + b"\0")
        w(average.tobytes())
        w(struct.pack("<Q", 0))                                     # no mask

    @staticmethod
    def frame_payload(recs, lines, pixels, line_begin=0, px_begin=0, timestamp_us=0, source_index=-1, flags=0) -> bytes:
        """Uncompressed payload of one frame (everything after the compression flag)."""
        n = len(recs)
        if n > 0xFFFF:
            raise ValueError("a PV frame holds at most 65535 objects")
        out = [struct.pack("<QHi", int(timestamp_us), n, int(source_index))]
        x0 = lines["x0"].astype(np.uint16); x1 = lines["x1"].astype(np.uint16); y = lines["y"]
        if len(lines) and int(x1.max()) >= 32768:
            raise ValueError("ShortHorizontalLine stores x in 15 bits")
        for r in recs:
            lo = int(r["line_off"]) - line_begin; nl = int(r["n_lines"])
            po = int(r["px_off"]) - px_begin; npx = int(r["n_pixels"])
            yy = y[lo:lo + nl]
            eol = np.zeros(nl, np.uint16)
            eol[:-1] = (yy[1:] != yy[:-1])                          # last line of a row; the blob's final line keeps 0
            pair = np.empty((nl, 2), "<u2")
            pair[:, 0] = x0[lo:lo + nl]
            pair[:, 1] = x1[lo:lo + nl] | (eol << 15)
            out.append(struct.pack("<HBH", int(yy[0]) if nl else 0, flags, nl))
            out.append(pair.tobytes())
            out.append(np.ascontiguousarray(pixels[po:po + npx]).tobytes())
        out.append(struct.pack("<H", 0))                            # no predictions
        return b"".join(out)

    def add_frame(self, recs, lines, pixels, line_begin=0, px_begin=0, timestamp_us=0, source_index=-1):
        self.index.append(self.f.tell())
        payload = self.frame_payload(recs, lines, pixels, line_begin, px_begin, timestamp_us, source_index)
        if self.compress:
            from .lzo1x import compress
            block = compress(payload)
            if len(block) + 8 < len(payload):                        # Frame::serialize keeps the smaller form (pv.cpp:741-763)
                self.f.write(b"\1" + struct.pack("<II", len(block), len(payload)) + block)
                self.compressed_frames += 1
                return
        self.f.write(b"\0")                                          # compression flag
        self.f.write(payload)

    def add_result(self, bs, i, timestamp_us=0, source_index=-1):
        """Append frame i of the last fetched batch of a trex_b200.BackgroundSubtraction."""
        info, recs, lines, px = bs.raw_result(i)
        self.add_frame(recs, lines, px, info.line_begin, info.px_begin, timestamp_us, source_index)

    def close(self):
        index_offset = self.f.tell()
        self.f.write(np.array(self.index, "<u8").tobytes())
        self.f.write(json.dumps({k: str(v) for k, v in self.metadata.items()}, separators=(",", ":")).encode() + b"\0")
        self.f.seek(self._num_frames_offset); self.f.write(struct.pack("<I", len(self.index)))
        self.f.seek(self._index_offset_offset); self.f.write(struct.pack("<Q", index_offset))
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
