"""Pins the oracle's pixel::find_outer_points (oracle/trex_oracle.c to_find_outer_points; SURVEY.md s8 row N4, first stage) on the REFERENCE'S OWN
CODE: commons/common/processing/PixelTree.cpp, compiled unmodified from the reference checkout (oracle/build_ref.py; the half of that file that
thresholds blobs only has to compile against oracle/ref_stubs/processing/pixeltree_standins.h and is never run).  Every outline of every blob,
in the reference's order, point for point -- the oracle works in bounding-box coordinates (what calculate_posture hands on), the reference returns
frame coordinates, so the box origin is subtracted (exact: integers from half-integers).  This also checks the oracle's one simplification: its
node set is "pixels with a missing 4-neighbour" instead of the reference's streaming three-row scan.
Runs wherever oracle/_ref/libref_posture.so exists or can be built; skipped otherwise."""
import ctypes as C

import numpy as np
import pytest

from oracle import build_ref, seg


@pytest.fixture(scope="module")
def ref():
    path = build_ref.build()
    if path is None:
        pytest.skip("no reference checkout and no prebuilt oracle/_ref/libref_posture.so")
    lib = C.CDLL(path)
    lib.ref_find_outer_points.restype = C.c_int64
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def ref_outlines(ref, lines):
    raw = np.ascontiguousarray(np.stack([lines["x0"], lines["x1"], lines["y"], np.zeros(len(lines), np.uint16)], 1).astype(np.uint16))
    npx = int((lines["x1"].astype(np.int64) - lines["x0"] + 1).sum())
    pts = np.zeros((4 * npx + 8, 2), np.float32)
    off = np.zeros(2 * npx + 8, np.int64)
    k = ref.ref_find_outer_points(_p(raw), C.c_int64(len(raw)), _p(pts), C.c_int64(len(pts)), _p(off), C.c_int64(len(off) - 1))
    assert k >= 0
    origin = np.array([lines["x0"].min(), lines["y"].min()], np.float32)
    return [pts[off[i]:off[i + 1]] - origin for i in range(k)]


def images():
    rng = np.random.default_rng(7)
    out = []
    for density in (0.35, 0.5, 0.62, 0.8):                      # dense random images: holes, diagonal contacts, single pixels, thin bridges
        for _ in range(3):
            out.append(((rng.random((40, 56)) < density) * 255).astype(np.uint8))
    yy, xx = np.mgrid[0:90, 0:160]
    for _ in range(6):                                          # smooth shapes with holes
        cx, cy, a, b = rng.uniform(50, 110), rng.uniform(30, 60), rng.uniform(20, 45), rng.uniform(8, 25)
        img = (((xx - cx) / a) ** 2 + ((yy - cy) / b) ** 2 <= 1)
        img &= ~((((xx - cx - 6) / (a * 0.3)) ** 2 + ((yy - cy) / (b * 0.4)) ** 2) <= 1)
        img |= (np.abs(yy - cy) <= 2) & (xx > cx) & (xx < cx + a + 25)
        out.append((img * 255).astype(np.uint8))
    one = np.zeros((5, 5), np.uint8); one[2, 2] = 255
    out.append(one)
    full = np.full((24, 300), 255, np.uint8); full[10:14, 100:200:7] = 0      # a wide blob with a row of holes
    out.append(full)
    return out


def test_find_outer_points_matches_the_compiled_reference(ref):
    n_blobs = n_loops = n_pts = 0
    for img in images():
        lab = seg.label_image(img)
        for b in range(len(lab)):
            lines, _ = lab.blob(b)
            mine = seg.find_outer_points(lines)
            want = ref_outlines(ref, lines)
            assert len(mine) == len(want), (img.shape, b, len(mine), len(want))
            for m, w in zip(mine, want):
                assert m.shape == w.shape and np.array_equal(m.view(np.uint32), np.ascontiguousarray(w).view(np.uint32)), (img.shape, b)
                n_pts += len(m)
            n_blobs += 1
            n_loops += len(mine)
    assert n_blobs > 300 and n_loops > n_blobs and n_pts > 20000
    print(f"{n_blobs} blobs, {n_loops} outlines, {n_pts} points: identical")
