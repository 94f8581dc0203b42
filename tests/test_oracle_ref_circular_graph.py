"""Pins the oracle's periodic:: restatements (oracle/trex_oracle.c: to_periodic_curvature, to_orientation_sum, to_eft, to_ieft, to_find_peaks -- the
arithmetic core of Outline::offset_to_middle) on the REFERENCE'S OWN CODE: commons/common/misc/CircularGraph.cpp, compiled
unmodified from the reference checkout by oracle/build_ref.py (its precompiled header is replaced by the stand-ins in oracle/ref_stubs/), called as
tracker/tracking/Outline.cpp:493-528 calls it.  Bit-exact: both sides are compiled without contraction and without fast-math.
Runs wherever oracle/_ref/libref_circulargraph.so exists or can be built (the reference checkout + g++); skipped otherwise."""
import ctypes as C

import numpy as np
import pytest

from oracle import build_ref, posture


@pytest.fixture(scope="module")
def ref():
    path = build_ref.build()
    if path is None:
        pytest.skip("no reference checkout and no prebuilt oracle/_ref/libref_circulargraph.so")
    lib = C.CDLL(path)
    lib.ref_orientation_sum.restype = C.c_float
    lib.ref_find_peaks.restype = C.c_int64
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def outlines():
    """Closed curves the way the tracker sees them: integer-lattice boundary walks of ellipses (find_outer_points), noisy smooth curves, sub-pixel
    resampled points, degenerate repeats."""
    rng = np.random.default_rng(11)
    out = []
    for k in range(24):
        n = int(rng.integers(12, 260))
        t = np.linspace(0, 2 * np.pi, n, endpoint=False)
        a, b = rng.uniform(4, 60), rng.uniform(3, 25)
        phi = rng.uniform(0, np.pi)
        x = a * np.cos(t) * np.cos(phi) - b * np.sin(t) * np.sin(phi) + rng.uniform(0, 500)
        y = a * np.cos(t) * np.sin(phi) + b * np.sin(t) * np.cos(phi) + rng.uniform(0, 500)
        pts = np.stack([x, y], 1)
        if k % 3 == 0:
            pts = np.round(pts)                                   # lattice points, with repeats
        elif k % 3 == 1:
            pts += rng.normal(0, 0.4, pts.shape)
        if k % 4 == 0:
            pts = pts[::-1]                                       # the other orientation
        out.append(np.ascontiguousarray(pts, np.float32))
    return out


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_curvature_orientation_eft_ieft_bit_exact(ref):
    for pts in outlines():
        n = len(pts)
        for r in (1, max(1, int(0.03 * n)), max(1, int(0.1 * n))):
            for absolute in (0, 1):
                got = np.zeros(n, np.float32)
                ref.ref_curvature(_p(pts), C.c_int64(n), r, absolute, _p(got))
                assert np.array_equal(bits(posture.curvature(pts, r, bool(absolute))), bits(got)), (n, r, absolute)
        s = ref.ref_orientation_sum(_p(pts), C.c_int64(n))
        assert np.float32(s).view(np.uint32) == np.float32(posture.orientation_sum(pts)).view(np.uint32)
        for order in (1, 3, 7):
            want = np.zeros((order, 4), np.float32)
            ref.ref_eft(_p(pts), C.c_int64(n), order, _p(want))
            coeffs = posture.eft(pts, order)
            assert np.array_equal(bits(coeffs), bits(want)), (n, order)
            for m in (n, 50):
                center = pts.mean(0).astype(np.float32)
                back = np.zeros((m, 2), np.float32)
                ref.ref_ieft(_p(want), order, C.c_int64(m), C.c_float(center[0]), C.c_float(center[1]), _p(back))
                assert np.array_equal(bits(posture.ieft(coeffs, m, (float(center[0]), float(center[1])))), bits(back)), (n, order, m)


def test_find_peaks_bit_exact(ref):
    rng = np.random.default_rng(5)
    curves = [posture.curvature(p, max(1, int(0.03 * len(p))), False) for p in outlines()]
    for k in range(12):                                           # synthetic profiles: plateaus, ties, sign changes at the wrap-around
        n = int(rng.integers(8, 200))
        v = np.round(rng.normal(0, 1, n), int(rng.integers(0, 3))).astype(np.float32)
        curves.append(np.roll(v, int(rng.integers(0, n))))
    checked = 0
    for v in curves:
        v = np.ascontiguousarray(v, np.float32)
        for broad in (0, 1):
            rec = np.zeros((len(v) + 1, 9), np.float32)
            n_ref = ref.ref_find_peaks(_p(v), C.c_int64(len(v)), broad, _p(rec), C.c_int64(len(rec)))
            got = posture.find_peaks(v, bool(broad))
            assert len(got) == n_ref, (len(v), broad)
            for i in range(n_ref):
                want = rec[i]
                mine = np.array([got[i]["x"], got[i]["y"], got[i]["width"], got[i]["integral"], got[i]["r0"], got[i]["r1"],
                                 got[i]["max_y_extrema"], got[i]["max_y"]], np.float32)
                assert np.array_equal(bits(mine), bits(want[:8])), (len(v), broad, i, mine, want)
                assert int(got[i]["n_pts"]) == int(want[8])
            checked += n_ref
    assert checked > 50
