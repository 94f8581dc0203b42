"""PV15 writer: round trip through the oracle's reader, and byte-for-byte equality of frame payloads with the
reference's own file (videos/test.pv stores its frames uncompressed) when the checkout is present."""
import os
import struct

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import seg
from trex_b200.background_subtraction import REC_DTYPE
from trex_b200.pv_writer import PVWriter

PV_PARAMS = seg.Params(detect_threshold=9, detect_size_filter=[(1, 10000)], cm_per_pixel=1.0)


def _recs_from(blobs):
    recs = np.zeros(len(blobs), REC_DTYPE)
    recs["line_off"] = blobs.line_off[:-1]; recs["n_lines"] = np.diff(blobs.line_off)
    recs["px_off"] = blobs.px_off[:-1]; recs["n_pixels"] = np.diff(blobs.px_off)
    return recs


def test_round_trip_without_lzo(tmp_path):
    g = np.load(os.path.join(GOLDEN, "testpv_golden.npz"))
    path = str(tmp_path / "out.pv")
    frames = []
    with PVWriter(path, 2304, 2304, g["average"], source="frames_%3d.jpg", name="t", metadata={"cm_per_pixel": 1, "meta_encoding": "gray"}) as w:
        for idx in (0, 100):
            b = seg.segment_frame(g[f"full{idx}_frame"], g["average"], PV_PARAMS)
            frames.append(b)
            w.add_frame(_recs_from(b), b.lines, b.pixels, timestamp_us=idx * 40000, source_index=idx)
    d = open(path, "rb").read()
    assert d[:5] == b"PV15\0" and d[5:10] == b"gray\0"
    # parse with a reader that needs no LZO: header fields, index, frames
    from oracle.pv15 import PV15
    pv = PV15.__new__(PV15)
    try:
        PV15.__init__(pv, path)
    except FileNotFoundError:
        pytest.skip("oracle/_ref/libminilzo.so not built (reader needs it only for compressed files)")
    assert (pv.width, pv.height, pv.num_frames, pv.encoding) == (2304, 2304, 2, "gray")
    assert np.array_equal(pv.average, g["average"])
    for i, b in enumerate(frames):
        assert pv.frame(i).as_list() == b.as_list()


@pytest.mark.skipif(not os.path.exists("/root/reference/videos/test.pv"), reason="reference checkout absent")
def test_frame_payload_bytes_equal_reference_file():
    cv2 = pytest.importorskip("cv2")
    from oracle.pv15 import PV15
    pv = PV15("/root/reference/videos/test.pv")
    for i in (0, 57, 199):
        fr = cv2.imread(f"/root/reference/videos/test_frames/frame_{i:03d}.jpg", cv2.IMREAD_UNCHANGED)
        b = seg.segment_frame(fr, pv.average, PV_PARAMS, seg.ORDER_REF_ABSORB)      # the order this file was written in
        pos = int(pv.index[i])
        assert pv.data[pos] == 0                                                      # stored uncompressed by TRex
        ts, n, src = struct.unpack_from("<QHi", pv.data, pos + 1)
        mine = PVWriter.frame_payload(_recs_from(b), b.lines, b.pixels, timestamp_us=ts, source_index=src)
        body = mine[:-2]                                                             # ours ends with "0 predictions"
        assert pv.data[pos + 1:pos + 1 + len(body)] == body


def _lzo_ref():
    import ctypes as C
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libminilzo.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libminilzo.so not built (needs the reference checkout)")
    L = C.CDLL(so)
    L.lzo1x_decompress_safe.argtypes = [C.c_char_p, C.c_ulong, C.c_char_p, C.POINTER(C.c_ulong), C.c_void_p]
    L.lzo1x_decompress_safe.restype = C.c_int

    def dec(block, n):
        out = C.create_string_buffer(n + 16); ol = C.c_ulong(n + 16)
        assert L.lzo1x_decompress_safe(block, len(block), out, C.byref(ol), None) == 0       # LZO_E_OK
        return out.raw[:ol.value]
    return dec


def test_lzo1x_encoder_round_trips_through_the_reference_decoder():
    """trex_b200.lzo1x.compress against minilzo's lzo1x_decompress_safe (the reader side of pv.cpp:315-336): every instruction
    form -- initial / extended literal runs, M2, M3, M4 (distances beyond 16 KiB), trailing literals -- on structured data."""
    import random
    from trex_b200.lzo1x import compress
    dec = _lzo_ref()
    rng = random.Random(1)
    blk = bytes(rng.randrange(256) for _ in range(40))
    cases = [b"", b"a", b"abc", b"abcd", b"a" * 1000, bytes(range(256)) * 4, b"x" * 17 + b"yz" + b"x" * 17,
             bytes(rng.randrange(256) for _ in range(5000)), bytes(rng.randrange(4) for _ in range(20000)),
             blk + bytes(rng.randrange(256) for _ in range(20000)) + blk + bytes(rng.randrange(256) for _ in range(30000)) + blk]
    for t in range(120):
        cases.append(bytes(rng.randrange(rng.choice([2, 3, 8, 64, 256])) for _ in range(rng.randrange(0, 2500))))
    for d in cases:
        assert dec(compress(d), len(d)) == d
    assert len(compress(b"a" * 1000)) < 20 and len(compress(b"hello world, " * 200)) < 80


def test_round_trip_with_lzo(tmp_path):
    """A file written with compress=True: frames are stored as LZO blocks and the oracle's reader (reference minilzo) returns
    the same blobs."""
    _lzo_ref()
    from oracle.pv15 import PV15
    g = np.load(os.path.join(GOLDEN, "testpv_golden.npz"))
    path = str(tmp_path / "lzo.pv")
    b = seg.segment_frame(g["full0_frame"], g["average"], PV_PARAMS)
    # a second, highly compressible frame: one blob of constant grey
    lines = np.array([(100, 400, y, 0) for y in range(50, 120)], seg.LINE_DTYPE)
    px = np.full(70 * 301, 37, np.uint8)
    recs2 = np.zeros(1, REC_DTYPE); recs2["n_lines"] = len(lines); recs2["n_pixels"] = len(px)
    with PVWriter(path, 2304, 2304, g["average"], compress=True) as w:
        w.add_frame(_recs_from(b), b.lines, b.pixels)
        w.add_frame(recs2, lines, px, timestamp_us=40000, source_index=1)
        assert w.compressed_frames >= 1
    pv = PV15(path)
    assert pv.num_frames == 2 and pv.data[int(pv.index[1])] == 1                 # stored compressed
    assert pv.frame(0).as_list() == b.as_list()
    f1 = pv.frame(1)
    assert len(f1) == 1 and np.array_equal(f1.lines, lines) and np.array_equal(f1.pixels, px)


def test_lzo1x_decoder_reads_minilzo_streams():
    """trex_b200.lzo1x.decompress on blocks produced by the reference's own compressor (lzo1x_1_compress, what TRex writes)."""
    import ctypes as C
    import random
    from trex_b200.lzo1x import compress, decompress
    _lzo_ref()
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libminilzo.so")
    L = C.CDLL(so)
    L.lzo1x_1_compress.argtypes = [C.c_char_p, C.c_ulong, C.c_char_p, C.POINTER(C.c_ulong), C.c_void_p]
    L.lzo1x_1_compress.restype = C.c_int
    wrk = C.create_string_buffer(1 << 17)
    rng = random.Random(5)
    blk = bytes(rng.randrange(256) for _ in range(64))
    cases = [b"", b"a", b"abc", b"a" * 1000, bytes(range(256)) * 8,
             blk + bytes(rng.randrange(256) for _ in range(30000)) + blk * 3 + bytes(rng.randrange(3) for _ in range(40000)) + blk]
    for _ in range(150):
        cases.append(bytes(rng.randrange(rng.choice([2, 3, 5, 16, 64, 256])) for _ in range(rng.randrange(0, 4000))))
    for d in cases:
        out = C.create_string_buffer(len(d) + len(d) // 16 + 67); ol = C.c_ulong(0)
        assert L.lzo1x_1_compress(d, len(d), out, C.byref(ol), wrk) == 0
        assert decompress(out.raw[:ol.value], len(d)) == d
        assert decompress(compress(d), len(d)) == d
    for junk in (b"\x00", b"\x11\x00", b"\xff" * 5, b"\x40\x00\x11\x00\x00"):
        with pytest.raises(ValueError):
            decompress(junk, 100)


def test_pv_reader_round_trip_and_reference_file(tmp_path):
    """trex_b200.pv_reader.PVReader: files from PVWriter (plain and LZO frames) and, when the checkout is present, the reference's
    own videos/test.pv against the oracle's reader."""
    from trex_b200.pv_reader import PVReader
    g = np.load(os.path.join(GOLDEN, "testpv_golden.npz"))
    b = seg.segment_frame(g["full0_frame"], g["average"], PV_PARAMS)
    lines = np.array([(100, 400, y, 0) for y in range(50, 120)], seg.LINE_DTYPE)
    px = np.full(70 * 301, 37, np.uint8)
    recs2 = np.zeros(1, REC_DTYPE); recs2["n_lines"] = len(lines); recs2["n_pixels"] = len(px)
    for compress_frames in (False, True):
        path = str(tmp_path / f"r{int(compress_frames)}.pv")
        with PVWriter(path, 2304, 2304, g["average"], name="rt", source="frames", metadata={"cm_per_pixel": 1}, compress=compress_frames) as w:
            w.add_frame(_recs_from(b), b.lines, b.pixels, timestamp_us=7, source_index=3)
            w.add_frame(recs2, lines, px, timestamp_us=40000, source_index=4)
        pv = PVReader(path)
        assert (pv.width, pv.height, len(pv), pv.encoding, pv.name, pv.source) == (2304, 2304, 2, "gray", "rt", "frames")
        assert np.array_equal(pv.average, g["average"]) and pv.mask is None and "cm_per_pixel" in pv.metadata
        f0, f1 = pv.frame(0), pv.frame(1)
        assert (f0.timestamp_us, f0.source_index, len(f0)) == (7, 3, len(b))
        assert np.array_equal(f0.lines, b.lines) and np.array_equal(f0.pixels, b.pixels)
        assert np.array_equal(f0.line_off, b.line_off) and np.array_equal(f0.px_off, b.px_off)
        assert np.array_equal(f1.lines, lines) and np.array_equal(f1.pixels, px)
    ref = "/root/reference/videos/test.pv"
    if os.path.exists(ref):
        from oracle.pv15 import PV15
        mine, theirs = PVReader(ref), PV15(ref)
        assert len(mine) == theirs.num_frames == 200 and np.array_equal(mine.average, theirs.average)
        for i in (0, 57, 199):
            a, c = mine.frame(i), theirs.frame(i)
            assert np.array_equal(a.lines, c.lines) and np.array_equal(a.pixels, c.pixels) and np.array_equal(a.line_off, c.line_off)
