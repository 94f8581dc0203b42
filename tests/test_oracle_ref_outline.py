"""Pins the oracle's posture chain on the REFERENCE'S OWN CODE: tracker/tracking/Outline.cpp, compiled unmodified from the
reference checkout by oracle/build_ref.py (stand-ins for TRex's precompiled header / settings cache in oracle/ref_stubs/) and called as
tracker/tracking/Posture.cpp:226-302 and Individual.cpp:507-522 call it.  Held bit for bit:
  Outline::resample                                   <-> oracle to_outline_resample
  Outline::calculate_midline  (smooth, the Fourier approximation, offset_to_middle with pointy / broad peaks, the pairing walk; segments,
                               tail / head indices AND the outline as the call leaves it)  <-> to_calculate_midline
  Midline::post_process       (with and without a movement direction)                     <-> to_midline_post_process
  Midline::normalize / fix_length                                                          <-> to_midline_normalize
The oracle restates the reference's three libm calls (atan2f, cosf, sinf) as the double function rounded to float, which differs from glibc 2.39's
atan2f by one ulp for ~17 % of the arguments (oracle/trex_oracle.c, "Third-party libm calls"); with `to_use_local_libm(1)` the oracle calls the local
libm like the compiled reference does, and then normalize is bit-exact too -- so the libm call is the ONLY difference, bounded here at 1e-4 px.
Runs wherever oracle/_ref/libref_posture.so exists or can be built; skipped otherwise."""
import ctypes as C

import numpy as np
import pytest

from oracle import build_ref, posture, seg


@pytest.fixture(scope="module")
def ref():
    path = build_ref.build()
    if path is None:
        pytest.skip("no reference checkout and no prebuilt oracle/_ref/libref_posture.so")
    lib = C.CDLL(path)
    lib.ref_outline_resample.restype = C.c_int64
    lib.ref_calculate_midline.restype = C.c_int64
    lib.ref_midline_normalize.restype = C.c_int64
    lib.ref_midline_post_process.restype = C.c_int
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def outlines(n=24, seed=3):
    """find_outer_points outlines of fish-like blobs (rotated ellipse + wavy tapering tail)."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        yy, xx = np.mgrid[0:70, 0:140]
        cx, cy = rng.uniform(40, 60), rng.uniform(28, 40)
        a, b = rng.uniform(18, 34), rng.uniform(6, 13)
        th = rng.uniform(-0.5, 0.5)
        xr = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)
        yr = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
        body = (xr / a) ** 2 + (yr / b) ** 2 <= 1
        tl = rng.uniform(20, 45)
        tail = (xr >= a * 0.7) & (xr <= a * 0.7 + tl) & (np.abs(yr + 0.15 * np.sin(xr / 9.0) * (xr - a * 0.7)) <= (a * 0.7 + tl - xr) * rng.uniform(0.12, 0.3))
        lines, _ = seg.label_image(((body | tail) * 255).astype(np.uint8)).blob(0)
        out.append(np.ascontiguousarray(seg.longest_outline(lines), np.float32))
    return out


def set_ref_settings(ref, P):
    ref.ref_outline_settings(C.c_float(P.outline_curvature_range_ratio), int(P.peak_mode), C.c_float(P.midline_walk_offset), int(P.outline_approximate),
                             int(P.outline_smooth_samples), int(P.outline_smooth_step), int(P.midline_start_with_head), int(P.midline_invert),
                             C.c_float(P.midline_stiff_percentage), C.c_uint32(P.midline_resolution))


def ref_midline(ref, resampled):
    pts = resampled.copy()
    n_after, t, h = C.c_int64(), C.c_int64(-1), C.c_int64(-1)
    segs = np.zeros((len(pts) + 8, 4), np.float32)
    ns = ref.ref_calculate_midline(_p(pts), C.c_int64(len(pts)), C.byref(n_after), _p(segs), C.c_int64(len(segs)), C.byref(t), C.byref(h))
    return ns, segs[:max(ns, 0)], int(t.value), int(h.value), pts[:n_after.value]


@pytest.mark.parametrize("distance", [1.0, 0.5, 2.5])
def test_resample_bit_exact(ref, distance):
    for ol in outlines(16):
        mine = seg.outline_resample(ol, distance)
        out = np.zeros((len(mine) + 64, 2), np.float32)
        n = ref.ref_outline_resample(_p(ol), C.c_int64(len(ol)), C.c_float(distance), _p(out), C.c_int64(len(out)))
        assert n == len(mine) and np.array_equal(bits(mine), bits(out[:n]))


SETTINGS = {
    "default": dict(),
    "broad": dict(peak_mode=1),
    "start_with_head": dict(midline_start_with_head=1),
    "invert": dict(midline_invert=1),
    "raw": dict(outline_approximate=0, outline_smooth_samples=0),
    "other": dict(outline_approximate=7, outline_smooth_samples=8, outline_smooth_step=2, outline_curvature_range_ratio=0.06, midline_walk_offset=0.05,
                  midline_stiff_percentage=0.3, midline_resolution=12),
}


@pytest.mark.parametrize("name", list(SETTINGS))
def test_midline_post_process_normalize_against_the_compiled_reference(ref, name):
    P = posture.default_params(**SETTINGS[name])
    set_ref_settings(ref, P)
    n_mid = n_pp = n_norm = n_norm_exact = 0
    worst = 0.0
    for ol in outlines():
        resampled = seg.outline_resample(ol, 1.0)
        ns, rsegs, rt, rh, rpts = ref_midline(ref, resampled)
        try:
            ms, mt, mh, mp = posture.calculate_midline(resampled, P)
        except posture.MidlineError as e:
            assert ns < 0 and np.array_equal(bits(e.points), bits(rpts))         # both refuse, and leave the same outline behind
            continue
        assert ns == len(ms) and np.array_equal(bits(ms), bits(rsegs)) and (mt, mh) == (rt, rh)
        assert np.array_equal(bits(mp), bits(rpts))
        n_mid += 1
        for move_dir in (None, (1.0, 0.0), (-0.6, 0.8)):
            rseg = ms.copy()
            t, h = C.c_int64(mt), C.c_int64(mh)
            md = None if move_dir is None else np.ascontiguousarray(move_dir, np.float32)
            rc = ref.ref_midline_post_process(_p(rseg), C.c_int64(len(rseg)), _p(md) if md is not None else None, C.byref(t), C.byref(h))
            try:
                pseg, pt, ph, pinv = posture.post_process(ms, P, move_dir, mt, mh)
            except IndexError:
                assert rc == -5
                continue
            assert rc == int(pinv) and np.array_equal(bits(pseg), bits(rseg)) and (pt, ph) == (t.value, h.value)
            n_pp += 1
            for fix_length in (-1.0, 40.0):
                out = np.zeros((int(P.midline_resolution) + 4, 4), np.float32)
                info = np.zeros(4, np.float32)
                nr = ref.ref_midline_normalize(_p(pseg), C.c_int64(len(pseg)), C.c_int64(pt), C.c_int64(ph), C.c_float(fix_length), _p(out), C.c_int64(len(out)), _p(info))
                for local in (1, 0):                                  # the local libm (as the compiled reference), then the oracle's own rounding
                    posture._lib().to_use_local_libm(local)
                    try:
                        got = posture.normalize(pseg, P, fix_length)
                    finally:
                        posture._lib().to_use_local_libm(0)
                    if got is None:
                        assert nr == 0
                        continue
                    gs, gl, ga, go = got
                    mine = np.array([gl, ga, go[0], go[1]], np.float32)
                    assert nr == len(gs)
                    if local:
                        assert np.array_equal(bits(gs), bits(out[:nr])) and np.array_equal(bits(mine), bits(info)), (name, fix_length)
                        n_norm_exact += 1
                    else:
                        d = max(float(np.abs(gs - out[:nr]).max()), float(np.abs(mine - info).max()))
                        worst = max(worst, d)
                        assert d < 1e-4, (name, fix_length, d)
                        n_norm += 1
    assert n_mid >= 20 and n_pp >= 3 * n_mid - 3 and n_norm_exact >= n_pp and n_norm >= n_pp
    print(f"{name}: {n_mid} midlines, {n_pp} post_process, {n_norm_exact} normalize bit-exact with the local libm; oracle rounding: worst |d| = {worst:.2e}")


def test_posture_crop_matrix_against_the_compiled_midline_transform(ref):
    """The 2 x 3 warp matrix of the `posture` / `legacy` crops: Midline::transform (Outline.cpp:1237-1256, compiled) composed with
    normalize_image's translate / scale / translate (FilterCache.cpp:47-63) through the reference's compiled gui::Transform -- bit-exact in double."""
    rng = np.random.default_rng(1)
    for k in range(2000):
        ang = np.float32(rng.uniform(-3.2, 3.2))
        off = rng.uniform(-60, 60, 2).astype(np.float32)
        length = np.float32(rng.uniform(5, 120))
        scale = np.float32(rng.choice([1.0, 0.5, 1.7]))
        legacy = k % 2
        M = np.zeros(6, np.float64)
        ref.ref_posture_matrix(C.c_float(ang), C.c_float(off[0]), C.c_float(off[1]), C.c_float(length), C.c_float(scale), legacy, 80, 80, _p(M))
        mine = np.asarray(posture.posture_matrix(float(ang), (float(off[0]), float(off[1])), float(length), (80, 80), float(scale), bool(legacy)), np.float64).reshape(-1)
        assert np.array_equal(mine.view(np.uint64), M.view(np.uint64)), (k, mine, M)
