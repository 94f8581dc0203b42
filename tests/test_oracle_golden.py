"""Pins the CPU oracle against the reference's own golden vectors (no GPU needed).

 * videos/test.pv blobs (tests/golden/testpv_golden.npz, made by tests/golden/make_golden.py)
 * Application/Tests/test_pixels.cpp:1381-1466 known-answer vector (pixels_golden.npz)
 * the reference's V118_3 class outputs (vi_golden.npz)
 * Application/Tests/test_matching.cpp:1556-1602 properties (one blob; render -> relabel idempotence)
When /root/reference is present (authoring container) all 200 fixture frames are checked as well.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import seg, vi

PV_PARAMS = seg.Params(detect_threshold=9, detect_size_filter=[(1, 10000)], cm_per_pixel=1.0)


def _blobs(g, prefix):
    return seg.Blobs(g[f"{prefix}_lines"], g[f"{prefix}_pixels"], g[f"{prefix}_line_off"], g[f"{prefix}_px_off"])


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "testpv_golden.npz"))


@pytest.mark.parametrize("idx", [0, 100])
def test_full_frames_match_test_pv(gold, idx):
    ref = _blobs(gold, f"full{idx}")
    got = seg.segment_frame(gold[f"full{idx}_frame"], gold["average"], PV_PARAMS, seg.ORDER_REF_ABSORB)
    assert len(got) == len(ref)
    assert got.as_list() == ref.as_list()          # exact file order with the absorb-size rule
    canon = seg.segment_frame(gold[f"full{idx}_frame"], gold["average"], PV_PARAMS)
    assert canon.as_set() == ref.as_set()
    first = [tuple(canon.blob(k)[0][["y", "x0"]][0]) for k in range(len(canon))]
    assert first == sorted(first)                  # canonical order = (y,x0) of first run


def test_windows_match_test_pv(gold):
    for (i, y0, x0) in gold["windows"]:
        fr = gold[f"win{i}_frame"]
        h, w = fr.shape
        ref = _blobs(gold, f"win{i}")
        got = seg.segment_frame(fr, gold["average"][y0:y0 + h, x0:x0 + w], PV_PARAMS)
        inner = set()
        for k in range(len(got)):
            l, p = got.blob(k)
            if l["x0"].min() > 0 and l["x1"].max() < w - 1 and l["y"].min() > 0 and l["y"].max() < h - 1:
                inner.add((l.tobytes(), p.tobytes()))
        assert inner == ref.as_set(), f"window of frame {i}"
        assert len(ref) > 0


@pytest.mark.skipif(not os.path.exists("/root/reference/videos/test.pv"), reason="reference checkout absent")
def test_all_200_frames_against_reference_checkout():
    cv2 = pytest.importorskip("cv2")
    from oracle.pv15 import PV15
    pv = PV15("/root/reference/videos/test.pv")
    total = 0
    for i in range(0, 200):
        fr = cv2.imread(f"/root/reference/videos/test_frames/frame_{i:03d}.jpg", cv2.IMREAD_UNCHANGED)
        ref = pv.frame(i)
        got = seg.segment_frame(fr, pv.average, PV_PARAMS, seg.ORDER_REF_ABSORB)
        assert got.as_list() == ref.as_list(), i
        total += len(ref)
    assert total == 39908


def test_image_from_lines_known_answer():
    g = np.load(os.path.join(GOLDEN, "pixels_golden.npz"))
    lines = np.zeros(2, seg.LINE_DTYPE)
    lines["x0"] = 0; lines["x1"] = 3; lines["y"] = [0, 1]
    rect, recount, mask, grey, diff = seg.image_from_lines(lines, g["pixels"], g["bg"], seg.DIFF_ABSOLUTE, int(g["threshold"]))
    assert list(rect) == [0, 0, 4, 2]
    assert recount == int(g["recount"])
    assert np.array_equal(mask, g["mask"])
    exp_diff = np.abs(g["bg"].astype(int) - g["pixels"].reshape(2, 4).astype(int)) * (g["mask"] > 0)
    assert np.array_equal(diff, exp_diff)
    assert np.array_equal(grey, g["pixels"].reshape(2, 4) * (g["mask"] > 0))


def _circle_rect_image():
    # test_matching.cpp:1556-1602: a filled circle overlapping a rectangle -> exactly one blob
    img = np.zeros((240, 320), np.uint8)
    yy, xx = np.mgrid[0:240, 0:320]
    img[(yy - 100) ** 2 + (xx - 120) ** 2 <= 50 ** 2] = 255
    img[90:160, 150:260] = 200
    return img


def test_ccl_single_blob_and_idempotence():
    img = _circle_rect_image()
    b = seg.label_image(img)
    assert len(b) == 1
    lines, px = b.blob(0)
    render = np.zeros_like(img)
    o = 0
    for l in lines:
        n = int(l["x1"]) - int(l["x0"]) + 1
        render[l["y"], l["x0"]:l["x1"] + 1] = px[o:o + n]; o += n
    assert np.array_equal(render, img)
    b2 = seg.label_image(render)
    assert b2.as_list() == b.as_list()


def test_edge_cases():
    P = seg.Params(detect_threshold=15, detect_size_filter=[])
    bg = np.full((32, 48), 100, np.uint8)
    fr = bg.copy()
    assert len(seg.segment_frame(fr, bg, P)) == 0                       # empty frame
    fr[31, 47] = 200; fr[0, 0] = 200                                     # corners, single pixels
    fr[10, 10:20] = 50; fr[11, 20] = 50                                  # diagonal touch -> one blob
    fr[20, 5:15] = 60; fr[20, 9] = 0                                     # grey 0 inside: splits (never foreground)
    fr[25, 5] = 116; fr[25, 7] = 115                                     # strict >: 116 is fg, 115 is not
    b = seg.segment_frame(fr, bg, P)
    got = sorted((int(b.blob(k)[0]["y"][0]), int(b.blob(k)[0]["x0"][0]), len(b.blob(k)[1])) for k in range(len(b)))
    assert got == [(0, 0, 1), (10, 10, 11), (20, 5, 4), (20, 10, 5), (25, 5, 1), (31, 47, 1)]
    assert seg.blob_id(b.blob(0)[0]) == (((0 + 0 + 1) // 2) << 19) | (0 << 6) | 1


def test_size_filter_half_open():
    bg = np.full((16, 64), 100, np.uint8)
    fr = bg.copy()
    fr[2, 2:12] = 10       # 10 px
    fr[6, 2:11] = 10       # 9 px
    fr[10, 2:22] = 10      # 20 px
    P = seg.Params(detect_threshold=15, detect_size_filter=[(10, 20)])
    b = seg.segment_frame(fr, bg, P)
    assert [len(b.blob(k)[1]) for k in range(len(b))] == [10]           # lo <= n < hi


def test_morphology_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    bg = np.full((96, 128), 120, np.uint8)
    fr = bg.copy()
    m = rng.random(fr.shape) < 0.08
    fr[m] = 30
    fr[40:60, 50:90] = 20; fr[48:52, 60:70] = 120
    for closing, k, dil in ((True, 3, 0), (True, 2, 0), (False, 3, 3), (False, 3, 2), (True, 1, 2)):
        P = seg.Params(detect_threshold=15, use_closing=closing, closing_size=k, dilation_size=dil)
        got = seg.generate_binary(fr, bg, P)
        d = cv2.absdiff(fr, bg)
        _, mask = cv2.threshold(d, 15, 255, cv2.THRESH_BINARY)
        if closing:
            el = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (2 * k + 1, 2 * k + 1), (k, k))
            mask = cv2.erode(cv2.dilate(mask, el), el)
        if dil > 0:
            mask = cv2.dilate(mask, np.ones((dil, dil), np.uint8))
        assert np.array_equal(got, cv2.bitwise_and(mask, fr)), (closing, k, dil)


def test_open_stage_matches_opencv_morphology_ex():
    """The optional n x n open of the threshold mask (BASELINE north_star's "2x2 morphological open"; no reference counterpart,
    SURVEY s0.5) is defined against cv2.morphologyEx(MORPH_OPEN, ones(n,n)) with OpenCV's default anchor."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(8)
    bg = np.full((90, 121), 120, np.uint8)
    fr = bg.copy()
    fr[rng.random(fr.shape) < 0.15] = 30
    fr[30:50, 40:80] = 20; fr[0:3, 0:3] = 10; fr[87:90, 118:121] = 10; fr[60, 5:100] = 10; fr[62:64, 5:100] = 10
    for n, closing in ((2, False), (3, False), (4, False), (2, True)):
        P = seg.Params(detect_threshold=15, open_size=n, use_closing=closing, closing_size=2)
        d = cv2.absdiff(fr, bg)
        _, mask = cv2.threshold(d, 15, 255, cv2.THRESH_BINARY)
        mask = cv2.morphologyEx(mask, cv2.MORPH_OPEN, np.ones((n, n), np.uint8))
        if closing:
            el = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (5, 5), (2, 2))
            mask = cv2.erode(cv2.dilate(mask, el), el)
        assert np.array_equal(seg.generate_binary(fr, bg, P), cv2.bitwise_and(mask, fr)), (n, closing)
    assert np.array_equal(seg.generate_binary(fr, bg, seg.Params(detect_threshold=15, open_size=1)),
                          seg.generate_binary(fr, bg, seg.Params(detect_threshold=15)))


def test_crop_geometry():
    bg = np.full((200, 300), 100, np.uint8)
    # small blob: centre pad, left/top get the larger half (FilterCache.cpp:187-196)
    lines = np.zeros(3, seg.LINE_DTYPE)
    lines["y"] = [50, 51, 52]; lines["x0"] = [100, 99, 100]; lines["x1"] = [104, 105, 103]
    px = np.arange(1, 1 + 5 + 7 + 4, dtype=np.uint8)
    c = seg.crop_blob(lines, px, bg, seg.DIFF_NONE)
    # bbox 7x3 at (99,50): left = 73 - 36 = 37, top = 77 - 38 = 39
    assert c.sum() == px.sum()
    assert np.array_equal(c[39, 38:43], px[0:5]) and np.array_equal(c[40, 37:44], px[5:12]) and np.array_equal(c[41, 38:42], px[12:16])
    cd = seg.crop_blob(lines, px, bg, seg.DIFF_ABSOLUTE)
    assert np.array_equal(cd[40, 37:44], 100 - px[5:12])
    # large blob: centre crop, start = d - d/2 (FilterCache.cpp:211-227)
    lines = np.zeros(1, seg.LINE_DTYPE); lines["y"] = 10; lines["x0"] = 20; lines["x1"] = 120   # 101 wide
    px = np.arange(101, dtype=np.uint8) + 1
    c = seg.crop_blob(lines, px, bg, seg.DIFF_NONE)
    assert np.array_equal(c[40], px[11:91])        # d = 21 -> start 11


def test_vi_nets_oracle_matches_reference_class_outputs():
    """V100 / V110 / V119 / V200 restated in oracle.vi against the outputs of the reference's own classes."""
    g = np.load(os.path.join(GOLDEN, "vi_nets_golden.npz"))
    for arch in ("v100", "v110", "v119", "v200"):
        for M, CI in ((12, 1), (9, 3)):
            tag = f"{arch}_m{M}c{CI}"
            sd = vi.scale_for_u8_inputs(vi.init_state_dict_arch(arch, M, CI, 80, 80, seed=0))
            assert vi.state_checksum(sd) == str(g[f"{tag}_checksum"])
            lg = vi.forward_logits_arch(arch, sd, g[f"{tag}_crops"])
            assert np.abs(lg - g[f"{tag}_logits"]).max() < 1e-5
            assert np.abs(vi.predict_arch(arch, sd, g[f"{tag}_crops"]) - g[f"{tag}_probs"]).max() < 1e-6
    assert [vi.arch_fc1_in(a) for a in ("v100", "v110", "v119", "v200")] == [10000, 10000, 3200, 512]


def test_vi_oracle_matches_reference_class_outputs():
    g = np.load(os.path.join(GOLDEN, "vi_golden.npz"))
    for tag, M in (("m100", 100), ("m8", 8)):
        sd = vi.scale_for_u8_inputs(vi.init_state_dict(M, 1, 80, 80, seed=0))
        assert vi.state_checksum(sd) == str(g[f"{tag}_checksum"])
        lg = vi.forward_logits(sd, g[f"{tag}_crops"])
        assert np.abs(lg - g[f"{tag}_logits"]).max() < 1e-5
        pr = vi.predict(sd, g[f"{tag}_crops"])
        assert np.abs(pr - g[f"{tag}_probs"]).max() < 1e-6
    assert [vi.batch_size_for(m) for m in (0, 8, 64, 65, 100, 128, 256, 1024)] == [64, 64, 64, 128, 128, 128, 128, 128]
    t = vi.transform_results(4, [0, 2, 3], {0: np.ones(3), 2: np.full(3, 2.0), 3: np.full(3, 3.0)}, 3)
    assert np.array_equal(t[1], [-1, -1, -1]) and np.array_equal(t[2], [2, 2, 2])


def test_averaging_matches_opencv():
    """oracle.average vs the OpenCV calls AveragingAccumulator makes (add in CV_32F, divide, convertTo; max; min)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(4)
    for n in (3, 10, 37):
        fr = rng.integers(0, 256, (n, 48, 64)).astype(np.uint8)
        acc = np.zeros((48, 64), np.float32)
        for f in fr:
            acc = cv2.add(acc, f.astype(np.float32))
        mean = np.clip(np.rint(cv2.divide(acc, float(n))), 0, 255).astype(np.uint8)     # convertTo(CV_8U) = cvRound + saturate
        assert np.array_equal(seg.average(fr, "mean"), mean)
        assert np.array_equal(seg.average(fr, "max"), fr.max(0)) and np.array_equal(seg.average(fr, "min"), fr.min(0))
    fr = rng.integers(100, 104, (25, 16, 16)).astype(np.uint8)
    mode = seg.average(fr, "mode")
    for y, x in ((0, 0), (5, 7), (15, 15)):
        c = np.bincount(fr[:, y, x], minlength=256)
        assert mode[y, x] == int(np.argmax(c))           # first maximum


def test_average_golden_from_test_pv():
    """The background image inside videos/test.pv (written by TRex, averaging_method=mode over 100 samples)
    is reproduced bit-exactly from the sampled test frames (committed window; all pixels when the reference is present)."""
    from trex_b200.averaging import sample_indices
    g = np.load(os.path.join(GOLDEN, "avg_golden.npz"))
    assert list(g["indices"]) == sample_indices(200, 100)
    assert np.array_equal(seg.average(g["frames"], "mode"), g["expected"])
    if os.path.exists("/root/reference/videos/test.pv"):
        cv2 = pytest.importorskip("cv2")
        from oracle.pv15 import PV15
        pv = PV15("/root/reference/videos/test.pv")
        rows = slice(1000, 1200)          # a 200-row band of the full 2304-wide frames keeps the test quick
        fr = np.stack([cv2.imread(f"/root/reference/videos/test_frames/frame_{i:03d}.jpg", cv2.IMREAD_UNCHANGED)[rows] for i in sample_indices(200, 100)])
        assert np.array_equal(seg.average(fr, "mode"), pv.average[rows])


def test_rethreshold_known_answers():
    """line_without_grid vectors of Application/Tests/test_pixels.cpp:981-1071 (gray): bg = 100, two lines of
    pixels 0,10,...,190, threshold 50, comparison >=, methods absolute / sign / none."""
    bg = np.full((10, 10), 100, np.uint8)
    lines = np.zeros(2, seg.LINE_DTYPE); lines["x0"] = 0; lines["x1"] = 9; lines["y"] = [0, 1]
    px = (np.arange(20) * 10).astype(np.uint8)
    parent = seg.Blobs(lines, px, np.array([0, 2], np.int64), np.array([0, 20], np.int64))

    def flat(b):
        return [(int(l["y"]), int(l["x0"]), int(l["x1"])) for l in b.lines], list(map(int, b.pixels))

    l, p = flat(seg.rethreshold(parent, bg, 50, seg.DIFF_ABSOLUTE))
    assert l == [(0, 0, 5), (1, 5, 9)] and p == [0, 10, 20, 30, 40, 50, 150, 160, 170, 180, 190]
    l, p = flat(seg.rethreshold(parent, bg, 50, seg.DIFF_SIGN))
    assert l == [(0, 0, 5)] and p == [0, 10, 20, 30, 40, 50]
    l, p = flat(seg.rethreshold(parent, bg, 50, seg.DIFF_NONE))
    assert l == [(0, 5, 9), (1, 0, 9)] and p == [50, 60, 70, 80, 90, 100, 110, 120, 130, 140, 150, 160, 170, 180, 190]
    # the absolute case splits into two blobs? rows 0 and 1 touch diagonally at x 5 -> one blob (8-connectivity)
    assert len(seg.rethreshold(parent, bg, 50, seg.DIFF_ABSOLUTE)) == 1


def test_box_mean_matches_opencv():
    """oracle.box_mean vs cv::boxFilter / cv::blur (the calls behind blur_difference and cv::adaptiveThreshold)."""
    import cv2
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (150, 217), dtype=np.uint8)
    img[50:90, 100:200] = 255; img[100:130] = 0
    assert np.array_equal(seg.box_mean(img, 25, "reflect101"), cv2.blur(img, (25, 25)))
    for k in (3, 7, 15, 17, 63, 255, 435, 1001):
        got = cv2.boxFilter(img, -1, (k, k), normalize=True, borderType=cv2.BORDER_REPLICATE | cv2.BORDER_ISOLATED)
        assert np.array_equal(seg.box_mean(img, k, "replicate"), got), k
    assert [seg.adaptive_neighbourhood(c, s) for c, s in ((1920, 2.0), (1920, 0.01), (100, 0.01), (640, 0.05))] == [3841, 19, 3, 33]


@pytest.mark.parametrize("absolute", [True, False])
def test_generate_binary_blur_difference_matches_opencv(absolute):
    """RawProcessing.cpp:371-387: difference -> THRESH_TOZERO -> blur 25x25 -> THRESH_BINARY -> mask & input."""
    import cv2
    rng = np.random.default_rng(5)
    bg = rng.integers(120, 140, (140, 200), dtype=np.uint8)
    fr = np.clip(bg.astype(int) + rng.integers(-12, 13, bg.shape), 1, 255).astype(np.uint8)
    fr[40:70, 60:120] = 30; fr[100:104, 20:24] = 250; fr[0:20, 180:200] = 60
    T = 9
    P = seg.Params(detect_threshold=T, detect_threshold_is_absolute=absolute, blur_difference=True, enable_difference=False)
    d = cv2.absdiff(fr, bg) if absolute else cv2.subtract(bg, fr)
    _, tz = cv2.threshold(d, T, 255, cv2.THRESH_TOZERO)
    _, m = cv2.threshold(cv2.blur(tz, (25, 25)), T, 255, cv2.THRESH_BINARY)
    exp = cv2.bitwise_and(m, fr)
    assert exp.any() and np.array_equal(seg.generate_binary(fr, bg, P), exp)


@pytest.mark.parametrize("T,scale,closing,dilation", [(9, 0.1, False, 0), (15, 2.0, False, 0), (-5, 0.05, False, 0), (9, 0.1, True, 0), (9, 0.06, False, -3), (12, 0.2, True, 2)])
def test_generate_binary_adaptive_threshold_matches_opencv(T, scale, closing, dilation):
    """use_adaptive_threshold (RawProcessing.cpp:427-434,487,526) followed by the same closing / dilation / erosion stages."""
    import cv2
    rng = np.random.default_rng(T + 100)
    bg = rng.integers(120, 140, (120, 170), dtype=np.uint8)
    fr = np.clip(bg.astype(int) + rng.integers(-6, 7, bg.shape), 1, 255).astype(np.uint8)
    fr[40:70, 60:120] = 30; fr[100:104, 20:24] = 250; fr[0:20, 150:170] = 60
    P = seg.Params(detect_threshold=T, use_adaptive_threshold=True, adaptive_threshold_scale=scale, use_closing=closing, closing_size=2,
                   dilation_size=dilation)
    d = cv2.absdiff(fr, bg)
    n = seg.adaptive_neighbourhood(fr.shape[1], scale)
    m = cv2.adaptiveThreshold(d, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY, n, -T)
    if T < 0:
        m = cv2.subtract(255, m)
    el = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (5, 5), (2, 2))
    if closing:
        m = cv2.erode(cv2.dilate(m, el), el)
    if dilation > 0:
        m = cv2.dilate(m, np.ones((dilation, dilation), np.uint8))
    elif dilation < 0:
        e = cv2.erode(m, np.ones((-dilation, -dilation), np.uint8))
        kept = np.where(e > 0, d, 0).astype(np.uint8)
        _, m = cv2.threshold(kept, abs(T), 255, cv2.THRESH_BINARY)
    exp = cv2.bitwise_and(m, fr)
    assert exp.any() and np.array_equal(seg.generate_binary(fr, bg, P), exp)


def test_resize_nearest_matches_opencv():
    """resize_image's default interpolation (C/misc/detail.h:465-469) as calculate_diff_image applies it for
    individual_image_scale != 1: oracle.resize_nearest vs cv::resize(INTER_NEAREST)."""
    import cv2
    rng = np.random.default_rng(0)
    for _ in range(200):
        w, h = int(rng.integers(1, 90)), int(rng.integers(1, 90))
        s = float(np.float32(rng.choice([0.5, 0.6, 0.75, 1.25, 1.5, 2.0, 0.33, 0.9, 1.1, 3.0])))
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        if int(np.rint(w * s)) == 0 or int(np.rint(h * s)) == 0:
            continue
        assert np.array_equal(seg.resize_nearest(img, s), cv2.resize(img, None, fx=s, fy=s, interpolation=cv2.INTER_NEAREST))
