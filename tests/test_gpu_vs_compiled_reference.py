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
