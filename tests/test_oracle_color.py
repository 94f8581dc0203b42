"""CPU tests of the oracle's colour path (no GPU): cv::cvtColor's fixed point against cv2, generate_binary on 3-channel
input against the OpenCV calls the reference makes (RawProcessing.cpp:355-358,529-534,581-589), and the rgb8
imageFromLines known answers of Application/Tests/test_pixels.cpp:1381-1466,1531-1583."""
import numpy as np
import pytest

from oracle import seg

cv2 = pytest.importorskip("cv2")


def test_bgr2gray_matches_opencv():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (257, 301, 3), dtype=np.uint8)
    assert np.array_equal(seg.bgr2gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
    img4 = rng.integers(0, 256, (64, 65, 4), dtype=np.uint8)
    assert np.array_equal(seg.bgr2gray(img4), cv2.cvtColor(img4, cv2.COLOR_BGRA2GRAY))
    grid = np.stack(np.meshgrid(np.arange(256), np.arange(0, 256, 3), np.arange(0, 256, 5), indexing="ij"), -1)
    grid = grid.reshape(1, -1, 3).astype(np.uint8)
    assert np.array_equal(seg.bgr2gray(grid), cv2.cvtColor(grid, cv2.COLOR_BGR2GRAY))


@pytest.mark.parametrize("T,absolute", [(15, True), (9, False), (40, True)])
def test_generate_binary_rgb8_matches_opencv_pipeline(T, absolute):
    rng = np.random.default_rng(T)
    bg = rng.integers(100, 160, (120, 160, 3), dtype=np.uint8)
    fr = np.clip(bg.astype(int) + rng.integers(-60, 60, bg.shape), 0, 255).astype(np.uint8)
    fr[5, 5] = (4, 0, 0); fr[6, 6] = 0                      # grey 0 but a non-zero channel / all zero
    P = seg.Params(detect_threshold=T, detect_threshold_is_absolute=absolute)
    out, gray = seg.generate_binary_color(fr, bg, P, encoding=seg.ENC_RGB8)
    g = cv2.cvtColor(fr, cv2.COLOR_BGR2GRAY); gb = cv2.cvtColor(bg, cv2.COLOR_BGR2GRAY)
    assert np.array_equal(gray, g)
    d = cv2.absdiff(g, gb) if absolute else cv2.subtract(gb, g)
    _, m = cv2.threshold(d, T, 255, cv2.THRESH_BINARY)
    exp = cv2.merge([cv2.bitwise_and(m, c) for c in cv2.split(fr)])
    assert np.array_equal(out, exp)
    # gray encoding of the same colour frame: cvtColor first, then the 1-channel path
    out1, _ = seg.generate_binary_color(fr, gb, P, encoding=seg.ENC_GRAY)
    assert np.array_equal(out1, seg.generate_binary(g, gb, P))
    out2, g2 = seg.generate_binary_color(fr, gb, P, encoding=seg.ENC_GRAY, color_channel=2)
    assert np.array_equal(g2, fr[..., 2]) and np.array_equal(out2, seg.generate_binary(fr[..., 2].copy(), gb, P))


def test_segment_frame_rgb8_consistency():
    rng = np.random.default_rng(3)
    bg = rng.integers(120, 140, (90, 130, 3), dtype=np.uint8)
    fr = bg.copy()
    fr[20:30, 40:70] = (30, 60, 90); fr[50:52, 10:12] = (0, 0, 0); fr[60:70, 100:110] = (250, 250, 250)
    P = seg.Params(detect_threshold=15, detect_size_filter=[(1, 100000)])
    b = seg.segment_frame_color(fr, bg, P, encoding=seg.ENC_RGB8)
    assert len(b) == 2                                            # the all-zero patch has no non-zero channel
    for k in range(len(b)):
        lines, px = b.blob(k)
        npx = int((lines["x1"].astype(int) - lines["x0"] + 1).sum())
        assert px.size == 3 * npx
        o = 0
        for l in lines:
            n = int(l["x1"]) - int(l["x0"]) + 1
            assert np.array_equal(px[o:o + 3 * n].reshape(n, 3), fr[l["y"], l["x0"]:l["x1"] + 1])
            o += 3 * n


def _vec():
    bg3 = np.array([[30, 50, 70, 90], [40, 60, 80, 100]], np.uint8)[..., None].repeat(3, -1)
    lines = np.zeros(2, seg.LINE_DTYPE); lines["x0"] = 0; lines["x1"] = 3; lines["y"] = [0, 1]
    vals = np.array([[25, 25, 25], [110, 110, 110], [80, 80, 80], [10, 200, 10],
                     [30, 30, 30], [95, 95, 95], [200, 200, 200], [100, 100, 100]], np.uint8)
    return bg3, lines, vals


def test_image_from_lines_rgb8_known_answer():
    """Application/Tests/test_pixels.cpp:1381-1466 (ImageFromLines.RGB8AbsoluteThresholdWithBackground)."""
    bg3, lines, vals = _vec()
    rect, n, mask, img, diff = seg.image_from_lines_rgb(lines, vals.reshape(-1), bg3, method=seg.DIFF_ABSOLUTE, base_threshold=25)
    assert tuple(rect) == (0, 0, 4, 2) and n == 4
    assert np.array_equal(mask, [[0, 255, 0, 255], [0, 255, 255, 0]])
    exp_img = np.array([[[0] * 3, [110] * 3, [0] * 3, [10, 200, 10]], [[0] * 3, [95] * 3, [200] * 3, [0] * 3]], np.uint8)
    exp_diff = np.array([[[0] * 3, [60] * 3, [0] * 3, [80, 110, 80]], [[0] * 3, [35] * 3, [120] * 3, [0] * 3]], np.uint8)
    assert np.array_equal(img, exp_img) and np.array_equal(diff, exp_diff)


def test_image_from_lines_rgb8_uses_all_channels():
    """test_pixels.cpp:1531-1583: one 200-valued channel over a background of 10 exceeds threshold 20."""
    bg3 = np.full((1, 1, 3), 10, np.uint8)
    lines = np.zeros(1, seg.LINE_DTYPE)
    for blob in ([200, 10, 10], [10, 200, 10], [10, 10, 200], [200, 200, 200]):
        rect, n, mask, _, _ = seg.image_from_lines_rgb(lines, np.array(blob, np.uint8), bg3, method=seg.DIFF_ABSOLUTE, base_threshold=20)
        assert tuple(rect) == (0, 0, 1, 1) and n == 1 and mask[0, 0] == 255


def test_tracker_grey_formula():
    """cmn::bgr2gray (Background.h:76-81) is float arithmetic, not OpenCV's fixed point: they differ on a few triples."""
    rng = np.random.default_rng(1)
    t = rng.integers(0, 256, (1, 200000, 3), dtype=np.uint8)
    a = seg.bgr2gray_tracker(t)
    exp = np.clip(t[..., 0].astype(np.float64) * 0.114 + t[..., 1] * 0.587 + t[..., 2] * 0.299 + 0.5, 0, 255).astype(np.uint8)
    assert np.array_equal(a, exp)
    frac = float((a != seg.bgr2gray(t)).mean())
    assert 0 < frac < 0.01


@pytest.mark.parametrize("last", [100, 90])
def test_rethreshold_rgb8_known_answer(last):
    """Application/Tests/test_pixels.cpp:1073-1166 (line_without_grid<rgb8 -> gray, absolute>) and :1289-1379
    (Blob::threshold on an rgb8 blob): threshold 25 against the grey image of the rgb background."""
    bg3, lines, vals = _vec()
    vals = vals.copy(); vals[7] = last
    bg_gray = cv2.cvtColor(bg3, cv2.COLOR_BGR2GRAY)                  # Background's _grey_image (Background.cpp:71-77)
    assert np.array_equal(bg_gray, seg.bgr2gray(bg3))
    blobs = seg.Blobs(lines, vals.reshape(-1).copy(), np.array([0, 2], np.int64), np.array([0, 24], np.int64))
    out = seg.rethreshold(blobs, bg_gray, 25, seg.DIFF_ABSOLUTE, rgb=True)
    assert len(out) == 1                                              # (0,1) (0,3) (1,1..2) are 8-connected
    got = [(int(l["y"]), int(l["x0"]), int(l["x1"])) for l in out.lines]
    assert got == [(0, 1, 1), (0, 3, 3), (1, 1, 2)]
    assert out.pixels.tolist() == [110, 110, 110, 10, 200, 10, 95, 95, 95, 200, 200, 200]
    # the grey twin of the same blob (pixels through cmn::bgr2gray) yields the same lines
    grey = seg.Blobs(lines, seg.bgr2gray_tracker(vals[None])[0].copy(), np.array([0, 2], np.int64), np.array([0, 8], np.int64))
    out_g = seg.rethreshold(grey, bg_gray, 25, seg.DIFF_ABSOLUTE)
    assert np.array_equal(out_g.lines, out.lines)
    assert np.array_equal(out_g.pixels, seg.bgr2gray_tracker(out.pixels.reshape(1, -1, 3))[0])


def test_r3g3b2_known_answers():
    """Application/Tests/test_pixels.cpp:629-795 (vec_to_r3g3b2, r3g3b2_to_vec, convert_to / convert_from)."""
    to = lambda *bgr: int(seg.convert_to_r3g3b2(np.array([[bgr]], np.uint8))[0, 0])
    assert to(255, 128, 64) == 0b11100010 and to(255, 128, 64, 255) == 0b11100010
    assert [to(255, 0, 0), to(0, 255, 0), to(0, 0, 255), to(255, 255, 255), to(0, 0, 0)] == [0b11000000, 0b00111000, 0b00000111, 0xFF, 0]
    back = seg.convert_from_r3g3b2(np.array([0b11100010, 0b11000000, 0b00111000, 0b00000111, 0xFF], np.uint8))
    assert back.tolist() == [[192, 128, 64], [192, 0, 0], [0, 224, 0], [0, 0, 224], [192, 224, 224]]
    img = np.array([[[255, 0, 0], [0, 255, 0]], [[0, 0, 255], [255, 255, 255]]], np.uint8)      # ImageConversionTest :745-767
    assert seg.convert_from_r3g3b2(seg.convert_to_r3g3b2(img)).tolist() == [[[192, 0, 0], [0, 224, 0]], [[0, 0, 224], [192, 224, 224]]]


def test_generate_binary_r3g3b2_matches_opencv_pipeline():
    """r3g3b2: BackgroundSubtraction::apply converts the frame to codes (:151-158); generate_binary then runs its
    1-channel path on the codes (absdiff / threshold / bitwise_and as for a grey image)."""
    rng = np.random.default_rng(4)
    bg3 = rng.integers(90, 170, (96, 128, 3), dtype=np.uint8)
    fr = np.clip(bg3.astype(int) + rng.integers(-80, 80, bg3.shape), 0, 255).astype(np.uint8)
    fr[3, 3] = (63, 31, 31)                                  # code 0: never foreground
    fr4 = np.dstack([fr, np.full(fr.shape[:2], 255, np.uint8)])
    bg = seg.convert_to_r3g3b2(bg3)
    for T, absolute in ((15, True), (30, False)):
        P = seg.Params(detect_threshold=T, detect_threshold_is_absolute=absolute, detect_size_filter=[(2, 100000)])
        out, codes = seg.generate_binary_color(fr, bg, P, encoding=seg.ENC_R3G3B2)
        exp_codes = ((fr[..., 0] // 64) << 6) | ((fr[..., 1] // 32) << 3) | (fr[..., 2] // 32)
        assert np.array_equal(codes, exp_codes)
        d = cv2.absdiff(codes, bg) if absolute else cv2.subtract(bg, codes)
        _, m = cv2.threshold(d, T, 255, cv2.THRESH_BINARY)
        assert np.array_equal(out, cv2.bitwise_and(m, codes))
        out4, _ = seg.generate_binary_color(fr4, bg, P, encoding=seg.ENC_R3G3B2)
        assert np.array_equal(out4, out)
        b = seg.segment_frame_color(fr, bg, P, encoding=seg.ENC_R3G3B2)
        b1 = seg.segment_frame(codes, bg, P)
        assert b.as_list() == b1.as_list() and len(b) > 0
        lines, px = b.blob(0)
        crop = seg.crop_blob_r3g3b2(lines, px, bg, seg.DIFF_ABSOLUTE)
        assert crop.shape == (80, 80, 3) and crop.any()


def test_rgb8_size_filter_counts_payload_bytes():
    """BackgroundSubtraction.cpp:247-259: num_pixels = pixels->size() -- the blob's payload BYTES (3 per pixel for rgb8,
    CPULabeling.cpp:302-311) -- so with detect_size_filter [10, 100) a 4-pixel rgb8 blob (12 bytes) is kept and a 34-pixel
    one (102 bytes) is dropped, while the gray encoding of the same frame does the opposite."""
    bg = np.full((40, 64, 3), 128, np.uint8)
    fr = bg.copy()
    fr[5, 8:12] = (20, 30, 40)            # 4 pixels
    fr[20, 10:44] = (20, 30, 40)          # 34 pixels
    fr[30, 10:20] = (20, 30, 40)          # 10 pixels: kept by both
    P = seg.Params(detect_threshold=15, detect_size_filter=[(10, 100)])
    rgb = seg.segment_frame_color(fr, bg, P, encoding=seg.ENC_RGB8)
    assert sorted(len(rgb.blob(k)[1]) // 3 for k in range(len(rgb))) == [4, 10]
    gray = seg.segment_frame_color(fr, seg.bgr2gray(bg), P, encoding=seg.ENC_GRAY)
    assert sorted(len(gray.blob(k)[1]) for k in range(len(gray))) == [10, 34]
