"""Pins the oracle on the REFERENCE'S OWN pv::Blob: commons/common/processing/PVBlob.{h,cpp} (+ BlobIdentity.cpp, bid.h), compiled unmodified into
oracle/_ref/libref_pvblob.so (oracle/build_ref.py build_pvblob):
  pv::Blob(lines, pixels, flags)  -> init() -> calculate_properties: bounds, centre, pixel count; blob id        <-> the GPU's tb_blob_rec fields / seg.blob_id
  calculate_moments() -> orientation(), centre of mass                                                            <-> seg.blob_orientation  (the `moments` crop's angle)
  recount(threshold, background) = raw_recount * SQR(cm_per_pixel), threshold 0, the cache                        <-> seg.blob_recount      (= tb_seg_recount)
Bit for bit, for labelled blobs of random images, ellipses, and blobs of more than 1000 runs -- where calculate_moments splits the runs into four packages
with separate float sums (PVBlob.cpp:118-204).  Runs wherever the library exists or can be built; skipped otherwise."""
import ctypes as C

import numpy as np
import pytest

from oracle import build_ref, seg
from test_oracle_ref_background import METHODS, noisy_frame, runs_of
from test_oracle_ref_labeling import _p


@pytest.fixture(scope="module")
def ref():
    path = build_ref.build_pvblob()
    if path is None:
        pytest.skip("no reference checkout and no prebuilt oracle/_ref/libref_pvblob.so")
    lib = C.CDLL(path)
    lib.ref_pvblob_properties.restype = C.c_uint32
    return lib


def shapes():
    """Run lists: labelled noise, rotated ellipses of many sizes, and tall blobs with 1001 ... 4500 runs."""
    rng = np.random.default_rng(5)
    out = []
    for d in (0.35, 0.55, 0.7):
        img = ((rng.random((70, 110)) < d) * 255).astype(np.uint8)
        B = seg.label_image(img)
        out += [B.blob(k)[0] for k in range(len(B))]
    for H, W, n in ((300, 400, 25), (5000, 700, 6)):
        yy, xx = np.mgrid[0:H, 0:W]
        for _ in range(n):
            cx, cy = rng.uniform(0.2 * W, 0.8 * W), rng.uniform(0.2 * H, 0.8 * H)
            a, b, th = rng.uniform(3, 0.45 * H), rng.uniform(2, 0.2 * W), rng.uniform(0, np.pi)
            u = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th); v = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
            m = ((u / a) ** 2 + (v / b) ** 2 < 1) & (rng.random((H, W)) < 0.97)          # a few holes: several runs per row
            B = seg.label_image(m.astype(np.uint8) * 255)
            k = max(range(len(B)), key=lambda i: len(B.blob(i)[0]))
            out.append(B.blob(k)[0])
    return out


def test_properties_moments_and_id(ref):
    n = n_big = 0
    for l in shapes():
        raw = runs_of(l)
        o = np.zeros(10, np.float32)
        bid = int(ref.ref_pvblob_properties(_p(raw), C.c_int64(len(raw)), _p(o)))
        angle, (cx, cy) = seg.blob_orientation(l)
        assert np.float32(angle).view(np.uint32) == o[0].view(np.uint32), (len(l), angle, float(o[0]))
        assert (np.float32(cx).view(np.uint32), np.float32(cy).view(np.uint32)) == (o[1].view(np.uint32), o[2].view(np.uint32)), (len(l), cx, cy, o[1:3])
        x0, y0, x1, y1 = int(l["x0"].min()), int(l["y"].min()), int(l["x1"].max()), int(l["y"].max())
        npx = int((l["x1"].astype(np.int64) - l["x0"] + 1).sum())
        assert list(o[3:8]) == [x0, y0, x1 - x0 + 1, y1 - y0 + 1, npx]
        assert (float(o[8]), float(o[9])) == (float(np.float32(x0) + np.float32(x1 - x0 + 1) * np.float32(0.5)), float(np.float32(y0) + np.float32(y1 - y0 + 1) * np.float32(0.5)))
        if x1 < 8192 and y0 < 8192:
            assert bid == seg.blob_id(l)
        n += 1
        n_big += int(len(l) > 1000)
    assert n > 150 and n_big >= 4


@pytest.mark.parametrize("method", [seg.DIFF_ABSOLUTE, seg.DIFF_SIGN, seg.DIFF_NONE])
@pytest.mark.parametrize("colour", [False, True])
def test_recount(ref, method, colour):
    ref.ref_background_settings(*METHODS[method], 2 if colour else 0)
    rng = np.random.default_rng(6)
    n = 0
    for _ in range(2):
        frame, bg = noisy_frame(rng, colour=colour)
        if colour:
            g = np.repeat(seg.bgr2gray(bg)[:, :, None], 3, axis=2)          # B = G = R background (tests/test_oracle_ref_background.py on colourful ones)
            frame = np.where(frame == bg, g, frame); bg = g
            blobs = seg.segment_frame_color(frame, bg, seg.Params(detect_threshold=10, detect_size_filter=[]), seg.ENC_RGB8)
            grey_bg = seg.bgr2gray(bg)
        else:
            blobs = seg.segment_frame(frame, bg, seg.Params(detect_threshold=10, detect_size_filter=[]))
            grey_bg = bg
        for b in range(len(blobs)):
            l, p = blobs.blob(b)
            p = np.ascontiguousarray(p, np.uint8)
            raw = runs_of(l)
            for T in (0, 12, 45):
                for cm in (1.0, 0.3):
                    o = np.zeros(3, np.float32)
                    ref.ref_pvblob_recount(_p(raw), C.c_int64(len(raw)), _p(p), C.c_int64(len(p)), 3 if colour else 1, _p(np.ascontiguousarray(bg)), bg.shape[1], bg.shape[0],
                                           3 if colour else 1, int(colour), T, C.c_float(cm), _p(o))
                    want = seg.blob_recount(l, p, grey_bg, T, method, cm, 3 if colour else 1)
                    assert o[0].view(np.uint32) == np.float32(want).view(np.uint32) and o[1].view(np.uint32) == o[0].view(np.uint32), (b, T, cm, float(o[0]), float(want))
                    n += 1
    ref.ref_background_settings(1, 1, 0)
    assert n > 100
