"""N1 as far as this image allows: INTEGRATION.md's drop-in shim EXECUTED through the reference's own interface.  tests/cpp/_dropin/libshim_dropin.so
(tests/build_dropin.py, built in the authoring container) holds the shim bodies of BackgroundSubtraction::Data::set / apply(std::vector<TileImage>&&) --
the first cpp block of INTEGRATION.md, verbatim -- compiled against the reference's real tracker/python/BackgroundSubtraction.h and linked with
libtrexb200.so.  The same C entry (oracle/ref_detect.cpp: set_background(Image::Ptr&&), one TileImage with promise and callback, apply, future.get())
drives (a) that library = the CUDA path and (b) oracle/_ref/libref_detect.so = the reference's own BackgroundSubtraction.cpp with the real OpenCV.
The pv::Frame of both receives the same objects (runs, pixel bytes, flags; canonical order on the GPU side), the encoding is set, the callback ran once.
Skipped where the prebuilt libraries or cv2 are missing."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def libs():
    pytest.importorskip("cv2")
    import build_dropin
    from oracle import build_ref
    from cv_bridge import Bridge
    shim, ref = build_dropin.build(), build_ref.build_detect()
    if shim is None or ref is None or not os.path.exists(build_dropin.OUT_B):
        pytest.skip("no prebuilt drop-in shim / compiled reference (built where the reference checkout is)")
    out = []
    for path in (shim, build_dropin.OUT_B, ref):
        lib = C.CDLL(path)
        lib.ref_background_subtraction_apply.restype = C.c_int64
        out.append((lib, Bridge(lib)))
    return out


def _run(lib, frame, bg, cap):
    from test_gpu_vs_compiled_reference import _p
    h, w, ch = frame.shape
    bg = np.ascontiguousarray(bg)
    lines = np.zeros((cap, 4), np.uint16); px = np.zeros(cap * 3, np.uint8)
    lo = np.zeros(cap + 1, np.int64); po = np.zeros(cap + 1, np.int64); fl = np.zeros(cap, np.uint8)
    enc, called = C.c_int32(-1), C.c_int32(0)
    k = lib.ref_background_subtraction_apply(_p(np.ascontiguousarray(frame)), h, w, ch, _p(bg), 1 if bg.ndim == 2 else bg.shape[2], _p(lines), C.c_int64(cap),
                                             _p(px), C.c_int64(len(px)), _p(lo), _p(po), _p(fl), C.c_int64(cap), C.byref(enc), C.byref(called))
    assert k >= 0 and called.value == 1, (k, called.value)
    return [(lines[lo[i]:lo[i + 1]].tobytes(), px[po[i]:po[i + 1]].tobytes(), int(fl[i])) for i in range(k)], enc.value


@pytest.mark.parametrize("which,encoding,size,n_blobs,kw,filt", [
    (0, "gray", (1080, 1920), 100, dict(), [(10.0, 100000.0)]),
    (1, "rgb8", (272, 480), 30, dict(detect_threshold=20, detect_threshold_is_absolute=0), [(4.0, 600.0), (900.0, 100000.0)]),
])
def test_shim_behind_the_reference_interface(libs, which, encoding, size, n_blobs, kw, filt):
    from oracle import seg
    from test_gpu_vs_compiled_reference import ENC, _configure
    from trex_b200.synthetic import BlobWorld, to_color
    (shim, _), (ref, bridge) = libs[which], libs[2]
    h, w = size
    world = BlobWorld(h=h, w=w, n_blobs=n_blobs, seed=31, margin=30)
    frames = to_color(world.frames(3), seed=31, channels=4)           # TileImage.images[0] is BGRA (the shim declares 4 channels)
    bg3 = to_color(world.bg, seed=32, channels=3)
    bg = bg3 if encoding == "rgb8" else seg.bgr2gray(bg3)
    for lib in (shim, ref):
        _configure(lib, encoding, filt, detect_batch_size=2, **kw)
    total = 0
    for f in range(len(frames)):
        got, enc_got = _run(shim, frames[f], bg, cap=h * w // 4)
        want, enc_want = _run(ref, frames[f], bg, cap=h * w // 4)
        assert not bridge.errors, bridge.errors
        assert enc_got == enc_want == ENC[encoding]
        assert len(got) == len(want) and sorted(got) == sorted(want), (f, len(got), len(want))
        total += len(got)
    assert total > 20
