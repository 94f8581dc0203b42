"""CPU tests of the posture ORACLE (N4: raw midline, Midline::post_process / normalize, posture crops): every piece of oracle/posture.py against an
independent numpy formulation of the same reference formula, and the whole chain on a shape whose midline is known by
construction.  Reference: T/tracking/Outline.cpp:330-452,454-718,768-868, C/misc/CircularGraph.cpp:12-606.
The reference has no test vectors for these functions; they are pinned on the reference's own compiled code in
tests/test_oracle_ref_outline.py / tests/test_oracle_ref_circular_graph.py."""
import numpy as np
import pytest

from oracle import posture, seg


def _fish():
    yy, xx = np.mgrid[0:60, 0:120]
    body = ((xx - 45) / 30.0) ** 2 + ((yy - 30) / 11.0) ** 2 <= 1
    tail = (xx >= 70) & (xx <= 110) & (np.abs(yy - 30) <= (110 - xx) * 0.22)
    img = ((body | tail) * 255).astype(np.uint8)
    lines, _ = seg.label_image(img).blob(0)
    return seg.outline_resample(seg.longest_outline(lines), 1.0)        # blob-relative: x0 = 15, y0 = 19


def test_fast_cos_is_the_reference_polynomial():
    xs = np.linspace(-20, 20, 4001, dtype=np.float32)
    got = np.array([posture.fast_cos(float(x)) for x in xs])
    assert np.abs(got - np.cos(xs.astype(np.float64))).max() < 1.2e-3          # the parabola approximation's error bound
    assert posture.fast_cos(0.0) == pytest.approx(1.0, abs=1e-6)


def test_smooth_is_a_circular_triangular_filter():
    rng = np.random.default_rng(0)
    pts = rng.random((50, 2)).astype(np.float32) * 40
    w = np.array([(4 - abs(i)) / 4 for i in range(-4, 5)], np.float64); w /= w.sum()
    exp = np.zeros((50, 2))
    for i in range(50):
        for k, j in enumerate(range(i - 4, i + 5)):
            exp[i] += pts[j % 50] * w[k]
    assert np.abs(posture.smooth(pts, 4, 1) - exp).max() < 1e-4
    assert np.array_equal(posture.smooth(pts[:4], 4, 1), pts[:4])               # L <= samples: unchanged (:384)
    w2 = np.array([(4 - abs(i)) / 4 for i in range(-4, 5, 2)], np.float64); w2 /= w2.sum()   # step 2: range 2, step_row 4
    exp2 = np.zeros((50, 2))
    for i in range(50):
        for k, j in enumerate(range(i - 4, i + 5, 2)):
            exp2[i] += pts[j % 50] * w2[k]
    assert np.abs(posture.smooth(pts, 2, 2) - exp2).max() < 1e-4


def test_curvature_formula():
    t = np.linspace(0, 2 * np.pi, 200, endpoint=False)
    pts = np.stack([30 + 20 * np.cos(t), 25 + 9 * np.sin(t)], 1).astype(np.float32)
    r = 6
    p1, p3 = np.roll(pts, r, 0).astype(np.float64), np.roll(pts, -r, 0).astype(np.float64)
    p2 = pts.astype(np.float64)
    cross = (p2[:, 0] - p1[:, 0]) * (p3[:, 1] - p2[:, 1]) - (p2[:, 1] - p1[:, 1]) * (p3[:, 0] - p2[:, 0])
    sq = lambda a, b: ((a - b) ** 2).sum(1)
    exp = 2 * cross / np.sqrt(sq(p1, p2) * sq(p2, p3) * sq(p1, p3))
    assert np.abs(posture.curvature(pts, r) - exp).max() < 1e-4
    assert np.abs(posture.curvature(pts, r, absolute=True) - np.abs(exp)).max() < 1e-4
    # an ellipse bends most at the ends of its major axis
    c = posture.curvature(pts, r, absolute=True)
    assert int(np.argmax(c)) in (0, 100, 199, 99, 101, 1)
    dup = pts.copy(); dup[5] = dup[5 + r]
    assert posture.curvature(dup, r)[5] == 0                                    # coincident points: left at 0 (:87,107)


def test_orientation_and_fourier_approximation():
    ol = _fish()
    s = posture.orientation_sum(ol)
    assert posture.orientation_sum(ol[::-1].copy()) * s < 0
    area2 = abs(np.sum(ol[:, 0] * np.roll(ol[:, 1], -1) - np.roll(ol[:, 0], -1) * ol[:, 1]))
    assert abs(abs(s) - area2) / area2 < 0.05                                   # the wrap-around term enters with swapped operands
    co = posture.eft(ol, 3)
    # float64 restatement of eft with exact trigonometry
    d = (np.roll(ol, -1, 0) - ol).astype(np.float64)[:-1]
    dt = np.sqrt((d ** 2).sum(1)) + 1e-10
    cum = np.concatenate([[0], np.cumsum(dt)]); T = cum[-1]; phi = 2 * np.pi * cum
    exp = np.zeros((3, 4))
    for n in (1, 2, 3):
        c, s_ = np.cos(phi * n / T), np.sin(phi * n / T)
        norm = T / (2 * np.pi ** 2) / n ** 2
        cn = ((d / dt[:, None]) * np.diff(c)[:, None]).sum(0) * norm
        sn = ((d / dt[:, None]) * np.diff(s_)[:, None]).sum(0) * norm
        exp[n - 1] = [cn[0], sn[0], cn[1], sn[1]]
    assert np.abs(co - exp).max() < 0.05                                        # fast::cos is accurate to ~1e-3
    back = posture.ieft(co, len(ol), ol.mean(0))
    # three harmonics keep the gross shape: every reconstructed point lies close to the outline
    dist = np.sqrt(((back[:, None, :] - ol[None, :, :]) ** 2).sum(2)).min(1)
    assert dist.max() < 8 and dist.mean() < 3


def test_find_peaks_on_a_two_bump_signal():
    x = np.arange(200)
    v = (np.exp(-((x - 50) / 6.0) ** 2) * 2.0 + np.exp(-((x - 140) / 15.0) ** 2) * 1.0 + 0.05).astype(np.float32)
    pk = posture.find_peaks(v)
    assert sorted(int(p) for p in pk["x"]) == [50, 140]
    hi, lo = pk[np.argmax(pk["y"])], pk[np.argmin(pk["y"])]
    assert hi["x"] == 50 and hi["r0"] < 50 < hi["r1"] and lo["r0"] < 140 < lo["r1"]
    assert hi["n_pts"] > 0 and lo["n_pts"] > hi["n_pts"]                        # the broad bump holds more points above half height
    assert len(posture.find_peaks(np.full(50, 3.0, np.float32))) == 0           # flat: no sign change of the difference


def test_midline_of_a_fish_shape():
    ol = _fish()
    segs, tail, head, pts = posture.calculate_midline(ol)
    assert tail == 0 and 0 < head < len(pts)
    # the pointy end (x ~ 95 in blob coordinates) is the tail, the blunt end (x ~ 0) the head; both on the axis y ~ 11.5
    assert pts[tail][0] > 80 and pts[head][0] < 10
    assert abs(pts[tail][1] - 11.5) < 1.5 and abs(pts[head][1] - 11.5) < 1.5
    assert 60 < len(segs) < len(pts) // 2 + 2
    assert np.abs(segs[:, 1] - 11.5).max() < 2.0                                # midline points stay on the symmetry axis
    assert np.all(np.diff(segs[:, 0]) < 0.5)                                    # and run from the tail towards the head
    assert segs[:, 2].max() < 24 and np.allclose(segs[:, 3], segs[:, 2] / 2, atol=1e-3)      # height <= body width, l_length = half
    with pytest.raises(ValueError):
        posture.calculate_midline(np.array([[0, 0], [1, 0], [1, 1]], np.float32))
    # peak_mode broad (:621-650): the tail is the middle of the broadest high-curvature stretch -- for this shape the blunt
    # head end, whose curvature plateau carries the larger integral; tail and head swap sides, the midline stays on the axis
    sb, tb, hb, pb = posture.calculate_midline(ol, posture.default_params(peak_mode=1))
    assert tb == 0 and len(sb) > 40 and np.abs(sb[:, 1] - 11.5).max() < 2.5
    assert abs(pb[tb][0] - pb[hb][0]) > 60                                      # the two ends of the body
    # midline_invert swaps the two indices (:712-713)
    _, t2, h2, _ = posture.calculate_midline(ol, posture.default_params(midline_invert=1))
    assert (t2, h2) == (head, tail)


def test_calculate_posture_retry_loop():
    """Posture.cpp:305-400: the threshold is raised in steps of 2 until a midline is found; a blob with a faint halo needs no
    retry, a blob that only yields tiny outlines ends with the first outline and no midline."""
    yy, xx = np.mgrid[0:80, 0:140]
    bg = np.full((80, 140), 150, np.uint8)
    fr = bg.copy()
    body = ((xx - 60) / 35.0) ** 2 + ((yy - 40) / 10.0) ** 2 <= 1
    fr[body] = 60
    P = seg.Params(detect_threshold=15, detect_size_filter=[])
    b = seg.segment_frame(fr, bg, P)
    lines, px = b.blob(0)
    r = posture.calculate_posture(lines, px, bg, track_posture_threshold=9, outline_resample=1.0)
    assert r["threshold"] == 9 and r["segments"] is not None and len(r["segments"]) > 20 and r["tail"] == 0
    # same result as the stages called by hand (the blob is its own thresholded sub-blob)
    ol = seg.outline_resample(seg.longest_outline(lines), 1.0)
    segs, tail, head, _ = posture.calculate_midline(ol)
    assert np.array_equal(r["segments"], segs) and (r["tail"], r["head"]) == (tail, head)
    # a 2-pixel blob: outlines too short for a midline at every threshold -> outline only
    fr2 = bg.copy(); fr2[10, 10:12] = 60
    l2, p2 = seg.segment_frame(fr2, bg, P).blob(0)
    r2 = posture.calculate_posture(l2, p2, bg, track_posture_threshold=9)
    assert r2["segments"] is None and len(r2["outline"]) > 0


def test_midline_lengths_against_the_references_own_export():
    """Loose external corroboration of the whole chain (re-threshold -> outline -> resample 0.5 -> smooth -> Fourier
    approximation -> curvature peaks -> pairing walk): the length of the raw midline of fish blobs from videos/test.pv next to
    the `midline_length` TRex itself exported for the same fish and frames (videos/compare_data_automatic/test_fish*.csv,
    written with videos/test.settings: track_threshold 12, sign difference, track_posture_threshold 9, outline_resample 0.5).
    The csv was written from a slightly different .pv (pixel counts differ by a few pixels) and holds the post-processed
    length rounded to an integer, so the agreement is statistical: tests/golden/make_golden.py::make_posture."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "posture_golden.npz"))
    n = int(g["count"])
    assert n >= 30
    ratios, close = [], 0
    for i in range(n):
        lines, px, bg = g[f"b{i}_lines"], g[f"b{i}_pixels"], g[f"b{i}_bg"]
        _, _, ml, npx = g[f"b{i}_csv"]
        out = posture.calculate_posture(lines, px, bg, track_posture_threshold=9, outline_resample=0.5, method=seg.DIFF_SIGN)
        assert out["segments"] is not None and out["threshold"] == 9
        s = out["segments"]
        length = float(np.linalg.norm(np.diff(s[:, :2], axis=0), axis=1).sum())
        ratios.append(length / ml)
        close += abs(length - ml) <= 1.5
        assert abs(len(px) - npx) <= 0.05 * npx                                  # the same fish (make_posture's filter)
    ratios = np.array(ratios)
    assert abs(ratios.mean() - 1) < 0.03 and ratios.std() < 0.05
    assert close >= 0.8 * n                                                      # within 1.5 px of TRex's integer for most fish


# ---- N4, third stage: Midline::post_process / normalize / transform and the posture crop (Outline.cpp:870-1456, FilterCache.cpp:21-115)

def _arc_midline(n=60, bend=0.6):
    """A midline along a circular arc, from the tail (index 0) to the head, like calculate_midline returns it; height = a fish profile."""
    t = np.linspace(0, 1, n)
    x, y = 80 * np.sin(bend * t) / bend, 80 * (1 - np.cos(bend * t)) / bend
    h = 12 * np.sin(np.pi * t) + 1
    return np.stack([x + 5, y + 7, h, h / 2], 1).astype(np.float32)


def test_post_process_keeps_the_order_and_stiffens_the_head_part():
    segs = _arc_midline()
    out, tail, head, inv = posture.post_process(segs, tail=0, head=30)
    assert (tail, head, inv) == (0, 30, False) and out.shape == segs.shape
    # default settings: reversed for the stiff-head pass, reversed back at the end -> same order, tail part untouched
    P = posture.default_params()
    n = len(segs)
    center = int(min(n - 1, round(n * P.midline_stiff_percentage) + 1))
    assert np.array_equal(out[: n - 1 - center], segs[: n - 1 - center])          # everything behind the stiff part
    assert np.array_equal(out[:, 2:], segs[:, 2:])                               # heights / l_length travel with their points
    # the stiff part keeps its segment lengths while it is straightened towards the body axis
    d_in = np.linalg.norm(np.diff(segs[:, :2], axis=0), axis=1)
    d_out = np.linalg.norm(np.diff(out[:, :2], axis=0), axis=1)
    assert np.allclose(d_in, d_out, atol=1e-3)
    head_dir_in = segs[-1, :2] - segs[-center, :2]; head_dir_out = out[-1, :2] - out[-center, :2]
    axis = segs[-center, :2] - segs[-center - 6, :2]
    ang = lambda a, b: np.arccos(np.dot(a, b) / np.linalg.norm(a) / np.linalg.norm(b))
    assert ang(head_dir_out, axis) < ang(head_dir_in, axis)                      # less bent than before
    # a movement direction pointing from the head to the tail inverts the midline (:925-945) and swaps the indices
    d = posture.post_process(segs, move_dir=(-1.0, 0.0), tail=0, head=30)
    assert d[3] is True and (d[1], d[2]) == (30, 0)
    assert np.allclose(d[0][0, :2], segs[-1, :2], atol=20) and not np.allclose(d[0][0, :2], segs[0, :2], atol=20)
    same = posture.post_process(segs, move_dir=(1.0, 0.0), tail=0, head=30)
    assert same[3] is False and np.array_equal(same[0], out)
    # midline_stiff_percentage = 0: only the two reversals
    P0 = posture.default_params(midline_stiff_percentage=0.0)
    assert np.array_equal(posture.post_process(segs, P0)[0], segs)
    # <= 2 segments: untouched (:896-904)
    assert np.array_equal(posture.post_process(segs[:2])[0], segs[:2])
    # a stiff part that reaches the end of the midline: the reference's segments().at(i + 1) throws
    with pytest.raises(IndexError):
        posture.post_process(segs[:12], posture.default_params(midline_stiff_percentage=0.95))


def test_normalize_resamples_by_arc_length_and_rotates_onto_the_x_axis():
    segs, _, _, _ = posture.post_process(_arc_midline())
    P = posture.default_params()
    pts, length, angle, off = posture.normalize(segs, P)
    assert pts.shape == (25, 4)
    xy = segs[:, :2].astype(np.float64)
    cum = np.concatenate([[0], np.cumsum(np.linalg.norm(np.diff(xy, axis=0), axis=1))])
    assert abs(length - cum[-1]) < 0.05                                          # chords of 24 steps on a gentle arc
    # independent restatement: 25 points at equal arc length, last = the head end, then Midline::real_point's inverse
    tgt = np.linspace(0, cum[-1], 25)
    exp = np.stack([np.interp(tgt, cum, xy[:, 0]), np.interp(tgt, cum, xy[:, 1])], 1)
    a = angle + np.pi                                                            # real_point (:1258-1266): rotate by angle + pi, add offset
    R = np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]])
    real = pts[:, :2].astype(np.float64) @ R.T + off
    assert np.abs(real - exp[::-1]).max() < 0.05                                 # normalised points run from the head (origin) to the tail
    assert np.array_equal(off, segs[-1, :2]) and np.all(pts[0, :2] == 0)
    step = np.linalg.norm(np.diff(pts[:, :2], axis=0), axis=1)
    assert np.allclose(step, length / 24, atol=2e-3)
    # the straight (stiffened) head part lies on the x axis, pointing to +x
    k = int(25 * P.midline_stiff_percentage)
    assert np.abs(pts[: k + 1, 1]).max() < 0.3 and pts[k, 0] > 0
    # heights: the reference weights them the other way round than the positions (:1346-1347: pos = s0 + line * percent, but
    # height = s0.height * percent + s1.height * (1 - percent)) -- restated as written
    j = np.clip(np.searchsorted(cum, tgt, side="right") - 1, 0, len(cum) - 2)
    f = (tgt - cum[j]) / (cum[j + 1] - cum[j])
    h = segs[:, 2].astype(np.float64)
    hexp = (h[j] * f + h[j + 1] * (1 - f))[::-1]
    assert np.abs(pts[1:-1, 2] - hexp[1:-1]).max() < 0.05
    # other resolutions; degenerate inputs return nullptr
    for res in (8, 50):
        q = posture.normalize(segs, posture.default_params(midline_resolution=res))
        assert q is not None and len(q[0]) == res and abs(q[1] - cum[-1]) < 0.5
    assert posture.normalize(segs[:1]) is None
    assert posture.normalize(np.repeat(segs[:1], 5, 0)) is None                  # zero length
    # fix_length (Individual::fixed_midline): steps of fix_length / resolution from the tail, extrapolated past the head
    f = posture.normalize(segs, P, fix_length=float(length) * 1.3)
    assert f is not None and len(f[0]) == 25
    fstep = np.linalg.norm(np.diff(f[0][:, :2], axis=0), axis=1)
    assert np.allclose(fstep, length * 1.3 / 25, rtol=0.02)


def test_atan2_restatement_against_the_local_libm():
    """The oracle (and the device) evaluate the reference's ::atan2f as (float)atan2(double) = the correctly rounded value (what glibc
    >= 2.41 returns).  The reference's own result depends on its platform's libm: this image's glibc 2.39 atan2f is faithfully, not
    correctly, rounded -- never more than one ulp away, equal for ~5 of 6 inputs."""
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.atan2f.restype = ctypes.c_float; libm.atan2f.argtypes = [ctypes.c_float, ctypes.c_float]
    rng = np.random.default_rng(5)
    ys, xs = (rng.standard_normal(4000) * 30).astype(np.float32), (rng.standard_normal(4000) * 30).astype(np.float32)
    ours = np.arctan2(ys.astype(np.float64), xs.astype(np.float64)).astype(np.float32)
    theirs = np.array([libm.atan2f(float(y), float(x)) for y, x in zip(ys, xs)], np.float32)
    ulp = np.abs(ours.view(np.int32).astype(np.int64) - theirs.view(np.int32).astype(np.int64))
    assert ulp.max() <= 1 and (ulp == 0).mean() > 0.75


def test_posture_matrix_and_crop_match_an_independent_restatement_and_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    yy, xx = np.mgrid[0:120, 0:200]
    th = 0.5
    u = (xx - 100) * np.cos(th) + (yy - 60) * np.sin(th); v = -(xx - 100) * np.sin(th) + (yy - 60) * np.cos(th)
    mask = (u / 30) ** 2 + (v / (8 * (1 - 0.7 * (u / 30))).clip(1)) ** 2 < 1
    bg = rng.integers(150, 220, (120, 200)).astype(np.uint8)
    fr = bg.copy(); fr[mask] = rng.integers(20, 90, int(mask.sum())).astype(np.uint8)
    b = seg.segment_frame(fr, bg, seg.Params(detect_threshold=15, detect_size_filter=[(10.0, 100000.0)]))
    lines, px = b.blob(0)
    ol = seg.outline_resample(seg.longest_outline(lines), 1.0)
    segs, tail, head, _ = posture.calculate_midline(ol)
    pp, _, _, _ = posture.post_process(segs, tail=tail, head=head)
    npts, length, angle, off = posture.normalize(pp)
    for legacy in (False, True):
        M = posture.posture_matrix(angle, off, length, legacy=legacy)
        # numpy restatement in double: T(size / 2) . S(1) . T(...) . T(-front = 0) . R(DEGREE(angle')) . T(-offset)
        T = lambda x, y: np.array([[1, 0, x], [0, 1, y], [0, 0, 1.0]])
        a = np.float32(np.float64(-np.float32(angle)) + (np.pi if legacy else np.pi * np.float64(np.float32(0.25))))
        deg = np.float32(a) * (np.float32(1.0) / np.float32(np.pi) * np.float32(180))
        rad = np.float64(deg) * 3.141592654 / 180.0
        R = np.array([[np.cos(rad), -np.sin(rad), 0], [np.sin(rad), np.cos(rad), 0], [0, 0, 1.0]])
        t = T(-np.float64(np.float32(length) * np.float32(0.5)), 0) if legacy else T(*([np.float64(np.float32(np.float64(np.float32(length)) * 0.4))] * 2))
        exp = T(40, 40) @ t @ R @ T(-np.float64(off[0]), -np.float64(off[1]))
        assert np.abs(M - exp[:2]).max() < 1e-9
        rect, _, _, _, diff = seg.image_from_lines(lines, px, bg, seg.DIFF_ABSOLUTE, 0)
        ref = cv2.warpAffine(diff, M, (80, 80), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
        got = posture.crop_blob_posture(lines, px, bg, seg.DIFF_ABSOLUTE, angle, off, length, legacy=legacy)
        assert np.array_equal(got, ref) and got.any()
    # posture: the head end of the midline lands at the canvas centre + 0.4 * length on both axes, the body along the diagonal
    M = posture.posture_matrix(angle, off, length)
    hx, hy = M @ np.array([off[0], off[1], 1.0])
    assert abs(hx - (40 + 0.4 * length)) < 1e-3 and abs(hy - (40 + 0.4 * length)) < 1e-3
    got = posture.crop_blob_posture(lines, px, bg, seg.DIFF_ABSOLUTE, angle, off, length)
    ys, xs = np.nonzero(got > 0)
    cov = np.cov(np.stack([xs, ys]))
    assert abs(0.5 * np.arctan2(2 * cov[0, 1], cov[0, 0] - cov[1, 1]) - np.pi / 4) < 0.3
    assert posture.crop_blob_posture(lines, px, bg, seg.DIFF_ABSOLUTE, angle, off, -1.0) is None      # invalid midline_length (:32-39)


def test_normalized_midline_length_against_the_references_own_export():
    """The corroboration of test_midline_lengths_against_the_references_own_export one stage further: TRex exports Midline::len() of
    the post-processed, normalised midline (an integer in the csv); post_process + normalize on the oracle's raw midlines gives it
    within the rounding for most fish."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "posture_golden.npz"))
    n = int(g["count"])
    diffs = []
    for i in range(n):
        lines, px, bg = g[f"b{i}_lines"], g[f"b{i}_pixels"], g[f"b{i}_bg"]
        ml = float(g[f"b{i}_csv"][2])
        out = posture.calculate_posture(lines, px, bg, track_posture_threshold=9, outline_resample=0.5, method=seg.DIFF_SIGN)
        pp, _, _, _ = posture.post_process(out["segments"], tail=out["tail"], head=out["head"])
        nm = posture.normalize(pp)
        assert nm is not None
        diffs.append(nm[1] - ml)
    diffs = np.array(diffs)
    print("normalised length - exported midline_length: mean %.3f sd %.3f, within 1 px: %d / %d" % (diffs.mean(), diffs.std(), (np.abs(diffs) <= 1).sum(), n))
    assert abs(diffs.mean()) < 1.0 and (np.abs(diffs) <= 1.5).mean() >= 0.8
