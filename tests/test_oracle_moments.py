"""CPU tests of the oracle's `moments` crop normalisation (no GPU): cv::warpAffine's fixed point against cv2, the
orientation against an independent numpy evaluation of calculate_moments, and the whole crop against the OpenCV call the
reference makes (FilterCache.cpp:21-115, 329-341)."""
import numpy as np
import pytest

from oracle import seg

cv2 = pytest.importorskip("cv2")


def test_warp_affine_matches_opencv():
    rng = np.random.default_rng(0)
    for trial in range(40):
        h, w = int(rng.integers(3, 120)), int(rng.integers(3, 120))
        src = rng.integers(0, 256, (h, w), dtype=np.uint8)
        ang = rng.uniform(0, 2 * np.pi)
        sc = rng.uniform(0.5, 2.0) if trial % 3 == 0 else 1.0
        c, s = sc * np.cos(ang), sc * np.sin(ang)
        M = np.array([[c, -s, 40 - (c * w / 2 - s * h / 2) + rng.uniform(-9, 9)], [s, c, 40 - (s * w / 2 + c * h / 2) + rng.uniform(-9, 9)]])
        ref = cv2.warpAffine(src, M, (80, 80), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
        assert np.array_equal(seg.warp_affine_u8(src, M), ref), trial
    # degenerate: identity, pure integer shift, singular matrix
    src = rng.integers(0, 256, (80, 80), dtype=np.uint8)
    for M in (np.array([[1., 0, 0], [0, 1, 0]]), np.array([[1., 0, 7], [0, 1, -3]]), np.zeros((2, 3))):
        assert np.array_equal(seg.warp_affine_u8(src, M), cv2.warpAffine(src, M, (80, 80), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT))


def _ellipse_blob(rng, h=200, w=260):
    from trex_b200.synthetic import BlobWorld
    world = BlobWorld(h=h, w=w, n_blobs=6, seed=int(rng.integers(0, 1 << 30)), margin=40)
    fr = world.frames(1)[0]
    b = seg.segment_frame(fr, world.bg, seg.Params(detect_threshold=15, detect_size_filter=[(50, 100000)]))
    return world, b


def test_orientation_matches_numpy_restatement():
    rng = np.random.default_rng(1)
    world, b = _ellipse_blob(rng)
    assert len(b) >= 3
    for k in range(len(b)):
        lines, _ = b.blob(k)
        ang, (cx, cy) = seg.blob_orientation(lines)
        # float32 accumulation in the same order
        f = np.float32
        m00 = m01 = m10 = f(0)
        for l in lines:
            for x in range(int(l["x0"]), int(l["x1"]) + 1):
                m00 = f(m00 + f(1)); m01 = f(m01 + f(int(l["y"]))); m10 = f(m10 + f(x))
        assert (cx, cy) == (float(f(m10 / m00)), float(f(m01 / m00)))
        mu = dict(mu00=f(0), mu02=f(0), mu11=f(0), mu20=f(0))
        for l in lines:
            vy = int(f(int(l["y"])) - f(cy)); vx = int(f(int(l["x0"])) - f(cx))
            for x in range(int(l["x0"]), int(l["x1"]) + 1):
                mu["mu00"] = f(mu["mu00"] + f(1)); mu["mu02"] = f(mu["mu02"] + f(vy * vy))
                mu["mu11"] = f(mu["mu11"] + f(f(vx) * f(vy))); mu["mu20"] = f(mu["mu20"] + f(vx * vx))
                vx += 1
        inv = f(f(1) / mu["mu00"])
        y, x = f(f(2) * f(mu["mu11"] * inv)), f(f(mu["mu20"] * inv) - f(mu["mu02"] * inv))
        # agrees with a double-precision atan2 up to the error of the reference's 3rd-order polynomial (fast_atan: 4.9e-3 at
        # |z| = 1, so 2.5e-3 after the factor 0.5), and exactly with the same polynomial evaluated in float32
        assert abs(ang - 0.5 * np.arctan2(float(y), float(x))) < 3e-3
        ay, ax = abs(y), abs(x)
        z = f(ay / ax) if ay < ax else f(ax / ay)
        pa = f(f(f(0.97239411) + f(f(f(-0.19194795) * z) * z)) * z)
        a = pa if ay < ax else f(np.float64(np.pi / 2) - np.float64(pa))
        if x < 0: a = f(np.float64(np.pi) - np.float64(a))
        if y < 0: a = f(-a)
        assert ang == float(f(np.float64(0.5) * np.float64(a)))


def test_moments_crop_matches_opencv_call():
    """normalize_image: cv::warpAffine(image, padded, t, 80x80, INTER_LINEAR, BORDER_CONSTANT) with t from the transform."""
    rng = np.random.default_rng(2)
    world, b = _ellipse_blob(rng)
    for k in range(len(b)):
        lines, px = b.blob(k)
        rect, _, mask, grey, diff = seg.image_from_lines(lines, px, world.bg, seg.DIFF_ABSOLUTE, 0)
        ang, _ = seg.blob_orientation(lines)
        M = seg.moments_matrix(ang, int(rect[2]), int(rect[3]))
        ref = cv2.warpAffine(diff, M, (80, 80), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
        got = seg.crop_blob_moments(lines, px, world.bg, seg.DIFF_ABSOLUTE)
        assert np.array_equal(got, ref), k
        assert got.any()
        # the rotated blob lies along the canvas diagonal direction fixed by the +pi/4 offset: its own orientation is ~ pi/4
        ys, xs = np.nonzero(got > 0)
        if len(xs) > 50:
            cov = np.cov(np.stack([xs, ys]))
            ang2 = 0.5 * np.arctan2(2 * cov[0, 1], cov[0, 0] - cov[1, 1])
            assert abs(abs(ang2) - np.pi / 4) < 0.35, (k, ang2)
