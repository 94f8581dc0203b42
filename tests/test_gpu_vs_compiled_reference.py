"""GPU parity WITHOUT the oracle in between: the CUDA path (trex_b200.BackgroundSubtraction through the C ABI) against the REFERENCE'S OWN
BackgroundSubtraction::set_background + apply(), compiled unmodified (tracker/python/BackgroundSubtraction.cpp, commons/common/processing/RawProcessing.cpp,
CPULabeling / Brototype / Source ...: oracle/_ref/libref_detect.so, built in the authoring container by oracle/build_ref.py build_detect) and run with the real
OpenCV (every cv:: call forwarded to cv2 by tests/cv_bridge.py).  Per frame the blob list the reference's pv::Frame receives equals the GPU's: the same set of
(runs, pixel bytes); the GPU emits them in canonical order, the reference in its labeling's order.  BASELINE configs[2]'s frame shape (1920 x 1080, 100 moving
blobs) and a small one; BGR frames under gray encoding, BGRA frames under rgb8; default settings, closing + dilation, a signed difference with two size ranges.
Skipped where cv2 or the prebuilt library is missing."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ENC = {"gray": 0, "r3g3b2": 1, "rgb8": 2}


@pytest.fixture(scope="module")
def ref():
    pytest.importorskip("cv2")
    from oracle import build_ref
    from cv_bridge import Bridge
    path = build_ref.build_detect()
    if path is None:
        pytest.skip("no prebuilt oracle/_ref/libref_detect.so and no reference checkout")
    lib = C.CDLL(path)
    lib.ref_background_subtraction_apply.restype = C.c_int64
    return lib, Bridge(lib)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _configure(lib, encoding, filt, **kw):
    base = dict(enable_difference=1, detect_threshold_is_absolute=1, detect_threshold=15, threshold_maximum=255, use_closing=0, closing_size=3,
                use_adaptive_threshold=0, adaptive_threshold_scale=2.0, dilation_size=0, image_invert=0, tags_enable=0, tags_equalize_hist=0, tags_threshold=15,
                cm_per_pixel=1.0, blur_difference=0)
    for k, v in {**base, **kw}.items():
        lib.ref_detect_setting(k.encode(), C.c_double(float(v)))
    lib.ref_detect_meta_encoding(ENC[encoding])
    lib.ref_detect_color_channel(-1)
    flat = np.array(filt, np.float64).reshape(-1)
    lib.ref_detect_size_filter(_p(flat) if len(flat) else None, len(filt))


def _reference_blobs(lib, frame, bg, cap):
    """[(runs bytes in the GPU's line layout {x0, x1, y, 0} u16, pixel bytes)] in the reference's order."""
    h, w, ch = frame.shape
    bg = np.ascontiguousarray(bg)
    lines = np.zeros((cap, 4), np.uint16); px = np.zeros(cap * 3, np.uint8)
    lo = np.zeros(cap + 1, np.int64); po = np.zeros(cap + 1, np.int64); fl = np.zeros(cap, np.uint8)
    enc, called = C.c_int32(-1), C.c_int32(0)
    k = lib.ref_background_subtraction_apply(_p(np.ascontiguousarray(frame)), h, w, ch, _p(bg), 1 if bg.ndim == 2 else bg.shape[2], _p(lines), C.c_int64(cap),
                                             _p(px), C.c_int64(len(px)), _p(lo), _p(po), _p(fl), C.c_int64(cap), C.byref(enc), C.byref(called))
    assert k >= 0 and called.value == 1, (k, called.value)
    return [(lines[lo[i]:lo[i + 1]].tobytes(), px[po[i]:po[i + 1]].tobytes()) for i in range(k)]


@pytest.mark.parametrize("size,n_blobs", [((1080, 1920), 100), ((272, 480), 30)])
@pytest.mark.parametrize("encoding,channels", [("gray", 3), ("rgb8", 4)])
@pytest.mark.parametrize("kw,filt", [(dict(), [(10.0, 100000.0)]),
                                     (dict(use_closing=1, closing_size=2, dilation_size=1, detect_threshold=20), [(1.0, 100000.0)]),
                                     (dict(detect_threshold_is_absolute=0, detect_threshold=25), [(4.0, 300.0), (600.0, 100000.0)])])
def test_cuda_path_equals_the_compiled_reference(ref, size, n_blobs, encoding, channels, kw, filt):
    import trex_b200
    from oracle import seg
    from trex_b200.synthetic import BlobWorld, to_color
    lib, bridge = ref
    h, w = size
    world = BlobWorld(h=h, w=w, n_blobs=n_blobs, seed=21, margin=30)
    frames = to_color(world.frames(2), seed=21, channels=channels)
    bg3 = to_color(world.bg, seed=22, channels=3)
    bg = bg3 if encoding == "rgb8" else seg.bgr2gray(bg3)      # what the reference's caller hands to set_background (a grey average for gray encoding)
    _configure(lib, encoding, filt, **kw)
    settings = trex_b200.DetectSettings(meta_encoding=encoding, detect_size_filter=list(filt),
                                        **{k: (bool(v) if isinstance(getattr(trex_b200.DetectSettings(), k), bool) else v) for k, v in kw.items()})
    bs = trex_b200.BackgroundSubtraction(bg, settings=settings, max_batch=2, channels=channels, max_runs_per_frame=h * w // 16, max_pixels_per_frame=h * w)
    got = bs.apply(frames)
    n = 0
    for f in range(len(frames)):
        want = _reference_blobs(lib, frames[f], bg, cap=h * w // 4)
        assert not bridge.errors, bridge.errors
        mine = [(b.lines.tobytes(), b.pixels.tobytes()) for b in got[f]]
        assert len(mine) == len(want), (f, len(mine), len(want))
        assert sorted(mine) == sorted(want), f
        n += len(mine)
    assert n > 10
    bs.deinit()


def test_posture_loop_equals_the_compiled_reference():
    """posture::calculate_posture (tracker/tracking/Posture.cpp:305-400 with PixelTree / CPULabeling / Outline / CircularGraph, compiled unmodified into
    oracle/_ref/libref_posture.so) against tb_seg_posture_thresholded, blob by blob: which of the three outcomes, the outline the result carries, the midline
    segments, tail and head -- bit for bit, including midlines found only after +2 retries.  The workload of tests/test_gpu_midline.py's loop test."""
    import trex_b200
    from oracle import build_ref, posture
    from test_oracle_ref_outline import set_ref_settings
    from test_oracle_ref_posture import graded_frame
    path = build_ref.build()
    if path is None:
        pytest.skip("no prebuilt oracle/_ref/libref_posture.so and no reference checkout")
    ref = C.CDLL(path)
    ref.ref_calculate_posture.restype = C.c_int64
    T0 = 12
    set_ref_settings(ref, posture.default_params())
    ref.ref_posture_settings(int(T0), C.c_float(1.0))
    ref.ref_background_settings(1, 1, 0)                    # track_threshold_is_absolute, track_background_subtraction, gray
    fr, bg = graded_frame(4)
    frames = np.stack([fr, np.roll(fr, 7, axis=1)])
    det = trex_b200.BackgroundSubtraction(bg, settings=trex_b200.DetectSettings(detect_threshold=10, detect_size_filter=[]), max_batch=2)
    pst = trex_b200.BackgroundSubtraction(bg, settings=trex_b200.DetectSettings(detect_threshold=0, detect_size_filter=[]), max_batch=2)
    got = det.apply(frames)
    rounds, res = det.posture_thresholded(pst, track_posture_threshold=T0, outline_resample=1.0, fetch=2)
    flat = [b for blobs in got for b in blobs]
    assert res["n_blobs"] == len(flat) > 20
    n_mid = n_outline = n_none = 0
    for k, b in enumerate(flat):
        so, ns, tail, head = (int(v) for v in res["midlines"][k])
        ro, rn = int(res["outlines"][k][2]), int(res["outlines"][k][3])
        raw = np.zeros((len(b.lines), 4), np.uint16); raw[:, 0], raw[:, 1], raw[:, 2] = b.lines["x0"], b.lines["x1"], b.lines["y"]
        p = np.ascontiguousarray(b.pixels, np.uint8)
        cap = 4 * len(p) + 64
        pts = np.zeros((cap, 2), np.float32); segs = np.zeros((cap, 4), np.float32)
        n_pts, t, h = C.c_int64(), C.c_int64(-1), C.c_int64(-1)
        r = ref.ref_calculate_posture(_p(raw), C.c_int64(len(raw)), _p(p), C.c_int64(len(p)), 1, _p(bg), bg.shape[1], bg.shape[0], 1, 0,
                                      _p(pts), C.c_int64(cap), C.byref(n_pts), _p(segs), C.c_int64(cap), C.byref(t), C.byref(h))
        if r == -2:                                         # "Cannot find valid posture"
            assert ns == 0 and rn == 0, k
            n_none += 1
            continue
        assert r >= -1, (k, r)
        mine = np.ascontiguousarray(res["points"][ro:ro + rn], np.float32)
        assert mine.shape == pts[:n_pts.value].shape and np.array_equal(mine.view(np.uint32), pts[:n_pts.value].view(np.uint32)), k
        if r == -1:                                         # an outline, no midline
            assert ns == 0, k
            n_outline += 1
        else:
            assert ns == r and (tail, head) == (t.value, h.value), (k, ns, r)
            assert np.array_equal(np.ascontiguousarray(res["segments"][so:so + ns], np.float32).view(np.uint32), segs[:r].view(np.uint32)), k
            n_mid += 1
    assert n_mid > 15 and n_outline + n_none >= 2 and rounds > 1
