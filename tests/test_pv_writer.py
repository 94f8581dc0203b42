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
