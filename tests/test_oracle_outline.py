"""CPU tests of the oracle's outline stage (N4, first stage): Outline::resample against the reference's own test vectors
(Application/Tests/test_outlines.cpp:53-94) and pixel::find_outer_points (C/processing/PixelTree.cpp:497-1130; no reference
vectors; pinned on the compiled PixelTree.cpp in tests/test_oracle_ref_pixeltree.py) on shapes whose outlines are known by construction."""
import numpy as np

from oracle import seg


def _lines(rows):
    return np.array([(x0, x1, y, 0) for (y, x0, x1) in rows], seg.LINE_DTYPE)


def test_resample_reference_vectors():
    sq = np.array([[0, 0], [10, 0], [10, 10], [0, 10]], np.float32)
    exp = np.array([[0, 0], [5, 0], [10, 0], [10, 5], [10, 10], [5, 10], [0, 10], [0, 5]], np.float32)
    got = seg.outline_resample(sq, 5.0)                                   # BasicFunctionality :53-61
    assert got.shape == exp.shape and np.abs(got - exp).max() <= 0.01    # compareOutlines tolerance :40
    assert len(seg.outline_resample(sq, 0.1)) > 100                       # VerySmallResamplingDistance :64-73
    assert len(seg.outline_resample(sq, 50.0)) < 3                        # VeryLargeResamplingDistance :75-84
    one = np.array([[0, 0]], np.float32)
    assert np.array_equal(seg.outline_resample(one, 5.0), one)            # SinglePointOutline :86-94
    assert np.array_equal(seg.outline_resample(sq, 0.0), sq)              # resampling_distance <= 0: unchanged (:727-728)


def test_find_outer_points_simple_shapes():
    # one pixel: its four side midpoints, blob on the right hand
    (o,) = seg.find_outer_points(_lines([(5, 7, 7)]))
    assert o.tolist() == [[0.0, 0.5], [0.5, 1.0], [1.0, 0.5], [0.5, 0.0]]
    # 3x3 ring: the outer outline (12 sides) first, then the hole (4 sides)
    outer, hole = seg.find_outer_points(_lines([(0, 0, 2), (1, 0, 0), (1, 2, 2), (2, 0, 2)]))
    assert len(outer) == 12 and hole.tolist() == [[1.5, 1.0], [2.0, 1.5], [1.5, 2.0], [1.0, 1.5]]
    assert np.array_equal(seg.longest_outline(_lines([(0, 0, 2), (1, 0, 0), (1, 2, 2), (2, 0, 2)])), outer)
    # two pixels touching by a corner are one 8-connected blob with ONE outline through the corner
    (o,) = seg.find_outer_points(_lines([(0, 0, 0), (1, 1, 1)]))
    assert len(o) == 8
    # every outline is closed with steps between neighbouring side midpoints (1, or sqrt(0.5) around a corner)
    rng = np.random.default_rng(3)
    img = (rng.random((40, 50)) < 0.6).astype(np.uint8) * 255
    b = seg.label_image(img)
    for k in range(len(b)):
        lines, px = b.blob(k)
        ols = seg.find_outer_points(lines)
        n_sides = 0
        for o in ols:
            d = np.linalg.norm(np.roll(o, -1, 0) - o, axis=1)
            assert np.all((np.abs(d - 1) < 1e-6) | (np.abs(d - np.sqrt(0.5)) < 1e-6))
            n_sides += len(o)
        # every missing 4-neighbour side of the blob lies on exactly one outline
        m = np.zeros((42, 52), bool)
        for l in lines:
            m[l["y"] + 1, l["x0"] + 1:l["x1"] + 2] = True
        exp = sum(int((m & ~np.roll(m, s, a)).sum()) for s, a in ((1, 0), (-1, 0), (1, 1), (-1, 1)))
        assert n_sides == exp


def test_outline_count_equals_one_plus_holes():
    """Independent structural check (scipy): an 8-connected blob has one outer outline plus one outline per hole, where holes
    are the 4-connected background regions that do not touch the border of the blob's padded bounding box."""
    ndimage = __import__("pytest").importorskip("scipy.ndimage")
    rng = np.random.default_rng(11)
    checked = 0
    blobs = []
    for dens in (0.35, 0.45, 0.55, 0.62):
        bl = seg.label_image((rng.random((70, 90)) < dens).astype(np.uint8) * 255)
        blobs += [bl.blob(k)[0] for k in range(len(bl))]
    for k, lines in enumerate(blobs):
        if len(lines) < 3:
            continue
        x0, y0 = int(lines["x0"].min()), int(lines["y"].min())
        w, h = int(lines["x1"].max()) - x0 + 1, int(lines["y"].max()) - y0 + 1
        m = np.zeros((h + 2, w + 2), bool)
        for l in lines:
            m[l["y"] - y0 + 1, l["x0"] - x0 + 1:l["x1"] - x0 + 2] = True
        lab, n = ndimage.label(~m)                          # 4-connectivity by default
        outside = lab[0, 0]
        holes = len(set(np.unique(lab)) - {0, outside})
        ols = seg.find_outer_points(lines)
        assert len(ols) == 1 + holes, k
        # the outer outline is the one that reaches the bounding box on all four sides
        outer = [o for o in ols if o[:, 0].min() == 0 and o[:, 1].min() == 0 and o[:, 0].max() == w and o[:, 1].max() == h]
        assert len(outer) == 1
        checked += 1
    assert checked > 20
