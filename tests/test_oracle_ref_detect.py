"""Pins the oracle's detection path (SURVEY.md s8 rows a-1 ... a-7: the function BASELINE.json's north_star names) on the REFERENCE'S OWN CODE run with
the REAL OpenCV:
  tracker/python/BackgroundSubtraction.cpp    BackgroundSubtraction::set_background + apply(std::vector<TileImage>&&)     <-> seg.segment_frame_color
  commons/common/processing/RawProcessing.cpp RawProcessing::generate_binary                                               <-> seg.generate_binary / generate_binary_color
  + CPULabeling / Brototype / Source / DLList / ListCache, tracker/core/SizeFilters.cpp, processing/Background.cpp
compiled unmodified into oracle/_ref/libref_detect.so (oracle/build_ref.py build_detect; stand-ins in oracle/ref_stubs/ and oracle/ref_stubs_detect/).
Both files are OpenCV call sequences; every cv:: call they make is forwarded to Python's cv2 (4.13) by tests/cv_bridge.py, so neither the control flow
nor the pixel arithmetic is restated anywhere between the reference and the comparison.  Compared: the binary image of every optional stage
(negative thresholds, threshold_maximum / inRange, signed difference, no difference, image_invert, closing with three element sizes, dilation, erosion
with its re-threshold, adaptive threshold, blur_difference) for gray and 3-channel input, and the blob list the pv::Frame receives -- runs, pixel bytes,
flags, in the reference's order -- for gray (BGR / BGRA frames, cvtColor or a colour channel), rgb8 and r3g3b2 encodings, size filters and cm_per_pixel.
Runs wherever cv2 and oracle/_ref/libref_detect.so exist (or the library can be built); skipped otherwise."""
import ctypes as C
import shutil

import numpy as np
import pytest

from oracle import build_ref, seg
from test_oracle_ref_labeling import _p, _unpack, oracle_blobs, same

cv2 = pytest.importorskip("cv2")
from cv_bridge import Bridge                                        # noqa: E402

DEFAULTS = dict(enable_difference=1, detect_threshold_is_absolute=1, detect_threshold=15, threshold_maximum=255, use_closing=0, closing_size=3,
                use_adaptive_threshold=0, adaptive_threshold_scale=2.0, dilation_size=0, image_invert=0, tags_enable=0, tags_equalize_hist=0, tags_threshold=15,
                cm_per_pixel=1.0)


def load(tmp_path_factory, blur_difference):
    """blur_difference (like the initial enable_difference) is a function-local static of generate_binary, read by the first call of a process
    (RawProcessing.cpp:265-267) and never refreshed: the other value needs its own copy of the library."""
    path = build_ref.build_detect()
    if path is None:
        pytest.skip("no reference checkout and no prebuilt oracle/_ref/libref_detect.so")
    if blur_difference:
        copy = str(tmp_path_factory.mktemp("ref") / "libref_detect_blur.so")
        shutil.copy(path, copy)
        path = copy
    lib = C.CDLL(path)
    lib.ref_background_subtraction_apply.restype = C.c_int64
    br = Bridge(lib)
    lib.ref_detect_setting(b"blur_difference", C.c_double(float(blur_difference)))
    configure(lib)
    return lib, br


@pytest.fixture(scope="module")
def ref(tmp_path_factory):
    return load(tmp_path_factory, 0)


@pytest.fixture(scope="module")
def ref_blur(tmp_path_factory):
    return load(tmp_path_factory, 1)


def configure(lib, **kw):
    for k, v in {**DEFAULTS, **kw}.items():
        lib.ref_detect_setting(k.encode(), C.c_double(float(v)))


def params(**kw):
    P = seg.Params(detect_size_filter=[])
    for k, v in kw.items():
        setattr(P, k, type(getattr(P, k))(v))
    return P


def scene(seed, H=96, W=128, colour=False):
    """A noisy background with dark and bright blobs, thin bridges, single pixels and a gradient: every stage changes some pixel."""
    rng = np.random.default_rng(seed)
    shape = (H, W, 3) if colour else (H, W)
    bg = rng.integers(90, 170, shape).astype(np.uint8)
    fr = np.clip(bg.astype(np.int32) + rng.integers(-6, 7, shape), 0, 255).astype(np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    for _ in range(9):
        cx, cy, a, b, th = rng.uniform(8, W - 8), rng.uniform(8, H - 8), rng.uniform(2, 14), rng.uniform(1, 6), rng.uniform(0, np.pi)
        u = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th); v = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
        m = (u / a) ** 2 + (v / b) ** 2 < 1
        lo, hi = ((0, 80) if rng.random() < 0.7 else (190, 256))
        fr[m] = rng.integers(lo, hi, (int(m.sum()), 3) if colour else int(m.sum())).astype(np.uint8)
    fr[10, 5:60] = 30; fr[40:44, 100] = 250; fr[70, 70] = 0; fr[0, 0:9] = 10; fr[H - 1, W - 4:] = 20
    return fr, bg


def ref_binary(lib, fr, bg):
    h, w = fr.shape[:2]
    ch = 1 if fr.ndim == 2 else 3
    out = np.zeros_like(fr)
    rc = lib.ref_generate_binary(_p(np.ascontiguousarray(fr)), h, w, ch, _p(np.ascontiguousarray(bg)), _p(out))
    assert rc == 0, rc
    return out


SWEEP = [dict(), dict(detect_threshold=40), dict(detect_threshold=-15), dict(detect_threshold=-40, use_closing=1, closing_size=2),
         dict(threshold_maximum=120), dict(threshold_maximum=60, detect_threshold=20, use_closing=1),
         dict(detect_threshold_is_absolute=0), dict(detect_threshold_is_absolute=0, detect_threshold=30, dilation_size=2),
         dict(enable_difference=0, detect_threshold=100), dict(enable_difference=0, detect_threshold=-100), dict(enable_difference=0, image_invert=1, detect_threshold=120),
         dict(image_invert=1), dict(image_invert=1, detect_threshold_is_absolute=0),
         dict(use_closing=1, closing_size=1), dict(use_closing=1, closing_size=3), dict(use_closing=1, closing_size=5, detect_threshold=25),
         dict(dilation_size=1), dict(dilation_size=3), dict(dilation_size=-2), dict(dilation_size=-3, detect_threshold=25),
         dict(use_closing=1, closing_size=2, dilation_size=2), dict(use_closing=1, closing_size=2, dilation_size=-2),
         dict(use_adaptive_threshold=1, adaptive_threshold_scale=0.1), dict(use_adaptive_threshold=1, adaptive_threshold_scale=0.03, detect_threshold=8),
         dict(use_adaptive_threshold=1, adaptive_threshold_scale=0.2, use_closing=1, closing_size=2), dict(use_adaptive_threshold=1, adaptive_threshold_scale=0.001, dilation_size=1)]


@pytest.mark.parametrize("kw", SWEEP, ids=lambda kw: ",".join(f"{k}={v}" for k, v in kw.items()) or "defaults")
def test_generate_binary_gray(ref, kw):
    lib, br = ref
    configure(lib, **kw)
    n_fg = 0
    for seed in (1, 2):
        fr, bg = scene(seed)
        got = ref_binary(lib, fr, bg)
        want = seg.generate_binary(fr, bg, params(**kw))
        assert not br.errors, br.errors
        assert np.array_equal(got, want), (kw, seed, int((got != want).sum()))
        n_fg += int((got != 0).sum())
    assert n_fg > 50
    configure(lib)


@pytest.mark.parametrize("kw", [dict(), dict(detect_threshold=30), dict(detect_threshold_is_absolute=0), dict(use_closing=1, closing_size=2), dict(dilation_size=2),
                                dict(threshold_maximum=100), dict(detect_threshold=-20)],
                         ids=lambda kw: ",".join(f"{k}={v}" for k, v in kw.items()) or "defaults")
def test_generate_binary_blur_difference(ref_blur, kw):
    lib, br = ref_blur
    configure(lib, **kw)
    for seed in (3, 4):
        fr, bg = scene(seed)
        got = ref_binary(lib, fr, bg)
        want = seg.generate_binary(fr, bg, params(blur_difference=1, **kw))
        assert not br.errors, br.errors
        assert "blur" in br.calls
        assert np.array_equal(got, want), (kw, seed, int((got != want).sum()))
    configure(lib)


@pytest.mark.parametrize("kw", [dict(), dict(detect_threshold=-25), dict(detect_threshold_is_absolute=0), dict(image_invert=1), dict(use_closing=1, closing_size=2, dilation_size=1),
                                dict(dilation_size=-2), dict(threshold_maximum=90)],
                         ids=lambda kw: ",".join(f"{k}={v}" for k, v in kw.items()) or "defaults")
def test_generate_binary_three_channels(ref, kw):
    """rgb8: input and average with three channels; the mask comes from the grey planes, the output keeps B, G, R under the mask (RawProcessing.cpp:363-372,562-599)."""
    lib, br = ref
    configure(lib, **kw)
    for seed in (5, 6):
        fr, bg = scene(seed, colour=True)
        got = ref_binary(lib, fr, bg)
        want, _ = seg.generate_binary_color(fr, bg, params(**kw), seg.ENC_RGB8)
        assert not br.errors, br.errors
        assert np.array_equal(got, want), (kw, seed, int((got != want).sum()))
    configure(lib)


def ref_apply(lib, fr, bg):
    h, w, ch = fr.shape
    bg = np.ascontiguousarray(bg)
    bg_ch = 1 if bg.ndim == 2 else bg.shape[2]
    n = h * w
    lines = np.zeros((n + 8, 4), np.uint16); px = np.zeros(n * 3 + 8, np.uint8)
    lo = np.zeros(n + 9, np.int64); po = np.zeros(n + 9, np.int64); fl = np.zeros(n + 8, np.uint8)
    enc, called = C.c_int32(-1), C.c_int32(0)
    k = lib.ref_background_subtraction_apply(_p(np.ascontiguousarray(fr)), h, w, ch, _p(bg), bg_ch, _p(lines), C.c_int64(len(lines)), _p(px), C.c_int64(len(px)),
                                             _p(lo), _p(po), _p(fl), C.c_int64(len(fl)), C.byref(enc), C.byref(called))
    assert k >= 0 and called.value == 1, (k, called.value)
    return _unpack(k, lines, px, lo, po, fl), enc.value


def set_filter(lib, ranges):
    flat = np.array(ranges, np.float64).reshape(-1)
    lib.ref_detect_size_filter(_p(flat) if len(flat) else None, len(ranges))


ENCODINGS = {seg.ENC_GRAY: 0, seg.ENC_R3G3B2: 1, seg.ENC_RGB8: 2}      # meta_encoding_t: gray, r3g3b2, rgb8, binary (processing/encoding.h:14)
IS_RGB, IS_R3G3B2 = 1 << 5, 1 << 6                                     # pv::Blob::Flags (PVBlob.h:138-145)


@pytest.mark.parametrize("encoding,channels,color_channel", [(seg.ENC_GRAY, 3, -1), (seg.ENC_GRAY, 4, -1), (seg.ENC_GRAY, 3, 1), (seg.ENC_GRAY, 4, 2), (seg.ENC_GRAY, 3, 7),
                                                             (seg.ENC_RGB8, 4, -1), (seg.ENC_R3G3B2, 3, -1), (seg.ENC_R3G3B2, 4, -1)])
@pytest.mark.parametrize("kw,filt,cm", [(dict(), [], 1.0), (dict(), [(10.0, 100000.0)], 1.0), (dict(detect_threshold=30), [(3.0, 40.0), (100.0, 400.0)], 1.0),
                                        (dict(use_closing=1, closing_size=2), [(0.5, 20.0)], 0.25), (dict(dilation_size=-2), [(4.0, 1000.0)], 0.5)])
def test_background_subtraction_apply(ref, encoding, channels, color_channel, kw, filt, cm):
    lib, br = ref
    configure(lib, cm_per_pixel=cm, **kw)
    lib.ref_detect_meta_encoding(ENCODINGS[encoding])
    lib.ref_detect_color_channel(color_channel)
    set_filter(lib, filt)
    P = params(cm_per_pixel=cm, **kw)
    P.detect_size_filter = list(filt)
    total = 0
    for seed in (7, 8):
        fr3, bg3 = scene(seed, colour=True)
        fr = fr3 if channels == 3 else np.concatenate([fr3, np.full(fr3.shape[:2] + (1,), 255, np.uint8)], axis=2)
        if encoding == seg.ENC_RGB8:
            bg = bg3
        elif encoding == seg.ENC_R3G3B2:
            bg = seg.convert_to_r3g3b2(bg3)
        elif 0 <= color_channel < 4:
            bg = np.ascontiguousarray(bg3[:, :, min(color_channel, 2)])
        else:
            bg = seg.bgr2gray(bg3)
        got, enc = ref_apply(lib, fr, bg)
        assert not br.errors, br.errors
        want = oracle_blobs(seg.segment_frame_color(fr, bg, P, encoding, color_channel, order=seg.ORDER_REF_LAZY))
        assert enc == ENCODINGS[encoding]
        assert same([(g[0], g[1]) for g in got], want), (seed, len(got), len(want))
        flag = IS_RGB if encoding == seg.ENC_RGB8 else (IS_R3G3B2 if encoding == seg.ENC_R3G3B2 else 0)
        assert all(g[2] == flag for g in got), sorted({g[2] for g in got})
        total += len(got)
    assert total > 3
    configure(lib); lib.ref_detect_meta_encoding(0); lib.ref_detect_color_channel(-1); set_filter(lib, [])


def test_apply_rejects_what_the_reference_rejects(ref):
    """A single-channel frame under gray encoding, a 3-channel frame under rgb8: "Invalid number of channels" -> the promise carries the exception
    (BackgroundSubtraction.cpp:157-186,324-328); the tile's callback still runs."""
    lib, br = ref
    configure(lib); set_filter(lib, [])
    fr3, bg3 = scene(9, colour=True)
    n = fr3.shape[0] * fr3.shape[1]
    lines = np.zeros((n, 4), np.uint16); px = np.zeros(3 * n, np.uint8); lo = np.zeros(n + 1, np.int64); po = np.zeros(n + 1, np.int64); fl = np.zeros(n, np.uint8)
    for enc, frame, bg in ((0, fr3[:, :, :1], seg.bgr2gray(bg3)), (2, fr3, bg3)):
        lib.ref_detect_meta_encoding(enc)
        e, called = C.c_int32(-1), C.c_int32(0)
        frame = np.ascontiguousarray(frame)
        k = lib.ref_background_subtraction_apply(_p(frame), frame.shape[0], frame.shape[1], frame.shape[2], _p(np.ascontiguousarray(bg)), 1 if bg.ndim == 2 else 3,
                                                 _p(lines), C.c_int64(n), _p(px), C.c_int64(3 * n), _p(lo), _p(po), _p(fl), C.c_int64(n), C.byref(e), C.byref(called))
        assert k == -1 and called.value == 1
    lib.ref_detect_meta_encoding(0)


@pytest.mark.parametrize("method", ["mean", "mode", "max", "min"])
@pytest.mark.parametrize("threaded", [0, 1])
def test_averaging_accumulator(ref, method, threaded):
    """commons/common/video/AveragingAccumulator.cpp compiled unmodified (row N2b): add / add_threaded over n frames, finalize.  cv::add on float matrices,
    cv::max / cv::min, cv::divide by the count and the float -> 8-bit convertTo are the real OpenCV's.  Against seg.average (= what tb_avg_* reproduces on
    the GPU): mean with its round-half-to-even, mode with ties (the smallest value wins), max, min; 1, 7 and 40 frames."""
    lib, br = ref
    rng = np.random.default_rng(13)
    for n in (1, 7, 40):
        frames = rng.integers(0, 256, (n, 60, 88)).astype(np.uint8)
        frames[:, :20] = rng.integers(100, 104, (n, 20, 88))           # few distinct values: ties for the mode, halves for the mean
        frames[:, 20:30] = (rng.integers(0, 2, (n, 10, 88)) * 255)      # saturated extremes
        out = np.zeros((60, 88), np.uint8)
        rc = lib.ref_average(_p(frames), n, 60, 88, 1, {"mean": 0, "mode": 1, "max": 2, "min": 3}[method], threaded, _p(out))
        assert rc == 0 and not br.errors, (rc, br.errors)
        want = seg.average(frames, method)
        assert np.array_equal(out, want), (method, n, int((out != want).sum()))


def test_generate_binary_random_setting_combinations(ref):
    """300 seeded random combinations of every setting generate_binary reads (thresholds from -200 to 255, threshold_maximum down to 0, difference on / off,
    signed / absolute, invert, closing 1 ... 6, dilation -4 ... 5, adaptive threshold with seven scales) on frames of random size: the compiled reference
    with the real OpenCV and the oracle agree on every pixel."""
    lib, br = ref
    rng = np.random.default_rng(99)
    for it in range(300):
        kw = dict(detect_threshold=int(rng.choice([-200, -60, -15, -1, 0, 1, 5, 15, 40, 120, 254, 255])),
                  threshold_maximum=int(rng.choice([255, 255, 255, 200, 90, 30, 0])),
                  enable_difference=int(rng.random() < 0.8), detect_threshold_is_absolute=int(rng.random() < 0.6),
                  image_invert=int(rng.random() < 0.3), use_closing=int(rng.random() < 0.4), closing_size=int(rng.integers(1, 7)),
                  dilation_size=int(rng.choice([0, 0, 0, 1, 2, 3, 5, -1, -2, -4])), use_adaptive_threshold=int(rng.random() < 0.25),
                  adaptive_threshold_scale=float(rng.choice([0.001, 0.02, 0.05, 0.1, 0.3, 1.0, 2.0])))
        configure(lib, **kw)
        fr, bg = scene(int(rng.integers(0, 1000)), H=int(rng.integers(72, 110)), W=int(rng.integers(102, 150)))
        got = ref_binary(lib, fr, bg)
        assert not br.errors, br.errors
        want = seg.generate_binary(fr, bg, params(**kw))
        assert np.array_equal(got, want), (it, kw, int((got != want).sum()))
    configure(lib)


def test_apply_random_configurations(ref):
    """120 seeded random configurations of BackgroundSubtraction::apply: encoding (gray / rgb8 / r3g3b2), BGR or BGRA frames, color_channel (none, 0 ... 3, out of range),
    thresholds of both signs, threshold_maximum, signed / absolute, invert, closing, dilation / erosion, cm_per_pixel, zero to two size ranges.  The frame receives the
    oracle's blob list in the reference's order -- except where EVERY surviving run lies in image row 0: there the reference's labeling returns nothing
    (Source::RowRef::from_index(0), DESIGN.md s6 deviation (2)), which this sweep meets twice."""
    lib, br = ref
    rng = np.random.default_rng(5)
    n_row0 = 0
    for it in range(120):
        enc = [seg.ENC_GRAY, seg.ENC_RGB8, seg.ENC_R3G3B2][int(rng.integers(0, 3))]
        ch = 4 if enc == seg.ENC_RGB8 else int(rng.choice([3, 4]))
        cc = -1 if enc != seg.ENC_GRAY else int(rng.choice([-1, -1, 0, 1, 2, 3 if ch == 4 else 2, 9]))
        kw = dict(detect_threshold=int(rng.choice([-40, -15, 5, 15, 40])), threshold_maximum=int(rng.choice([255, 255, 120])),
                  detect_threshold_is_absolute=int(rng.random() < 0.6), image_invert=int(rng.random() < 0.2), use_closing=int(rng.random() < 0.3), closing_size=int(rng.integers(1, 4)),
                  dilation_size=int(rng.choice([0, 0, 1, 2, -2])))
        cm = float(rng.choice([1.0, 0.5, 0.13]))
        nf = int(rng.integers(0, 3))
        filt = sorted([(float(a), float(a + b)) for a, b in zip(rng.uniform(0, 50, nf), rng.uniform(1, 400, nf))])
        configure(lib, cm_per_pixel=cm, **kw)
        lib.ref_detect_meta_encoding(ENCODINGS[enc]); lib.ref_detect_color_channel(cc); set_filter(lib, filt)
        P = params(cm_per_pixel=cm, **kw)
        P.detect_size_filter = list(filt)
        fr3, bg3 = scene(int(rng.integers(0, 1000)), colour=True)
        fr = fr3 if ch == 3 else np.concatenate([fr3, np.full(fr3.shape[:2] + (1,), 255, np.uint8)], axis=2)
        if enc == seg.ENC_RGB8:
            bg = bg3
        elif enc == seg.ENC_R3G3B2:
            bg = seg.convert_to_r3g3b2(bg3)
        elif 0 <= cc < 4:
            bg = np.ascontiguousarray(bg3[:, :, min(cc, 2)])
        else:
            bg = seg.bgr2gray(bg3)
        got, _ = ref_apply(lib, fr, bg)
        assert not br.errors, br.errors
        B = seg.segment_frame_color(fr, bg, P, enc, cc, order=seg.ORDER_REF_LAZY)
        unfiltered = seg.segment_frame_color(fr, bg, params(cm_per_pixel=cm, **kw), enc, cc)
        if len(unfiltered) and all(int(unfiltered.blob(k)[0]["y"].max()) == 0 for k in range(len(unfiltered))):
            assert got == [], (it, len(got))
            n_row0 += 1
            continue
        assert same([(g[0], g[1]) for g in got], oracle_blobs(B)), (it, enc, ch, cc, kw, cm, filt, len(got), len(B))
    configure(lib); lib.ref_detect_meta_encoding(0); lib.ref_detect_color_channel(-1); set_filter(lib, [])
    assert n_row0 <= 4
