"""Pins the oracle's posture::calculate_posture (oracle/posture.py calculate_posture; SURVEY.md s8 row N4: the threshold loop) on the REFERENCE'S OWN
CODE: tracker/tracking/Posture.cpp:305-400 compiled unmodified together with everything it calls -- pixel::threshold_get_biggest_blob (PixelTree.cpp over
Background.{h,cpp} and CPULabeling), pixel::find_outer_points, Outline::resample, Outline::calculate_midline -- by oracle/build_ref.py.  Per blob: which
of the three outcomes (midline / outline only / "Cannot find valid posture"), the outline the result carries, the midline segments, tail / head --
bit for bit.  The workload is the one of tests/test_gpu_midline.py::test_posture_of_thresholded_blobs_with_retries (graded blobs whose thresholded
sub-blobs shrink round by round, two-core blobs, tiny blobs), which holds the GPU loop to the same oracle.
Runs wherever oracle/_ref/libref_posture.so exists or can be built; skipped otherwise."""
import ctypes as C

import numpy as np
import pytest

from oracle import build_ref, posture, seg
from test_oracle_ref_labeling import _p
from test_oracle_ref_outline import set_ref_settings


@pytest.fixture(scope="module")
def ref():
    path = build_ref.build()
    if path is None:
        pytest.skip("no reference checkout and no prebuilt oracle/_ref/libref_posture.so")
    lib = C.CDLL(path)
    lib.ref_calculate_posture.restype = C.c_int64
    return lib


def graded_frame(seed, n=14, H=160, W=256):
    rng = np.random.default_rng(seed)
    bg = np.full((H, W), 200, np.uint8)
    fr = bg.copy()
    yy, xx = np.mgrid[0:H, 0:W]
    for k in range(n):                                     # elongated blobs with a darkness gradient, some with two dark cores
        cx, cy, a, b, th = rng.uniform(30, W - 30), rng.uniform(20, H - 20), rng.uniform(6, 22), rng.uniform(2, 6), rng.uniform(0, np.pi)
        u = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th); v = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
        r2 = (u / a) ** 2 + (v / b) ** 2
        m = r2 < 1
        depth = rng.uniform(20, 120)
        core = np.where(k % 5 == 0, np.minimum(((u - a / 2) / (a / 3)) ** 2 + (v / b) ** 2, ((u + a / 2) / (a / 3)) ** 2 + (v / b) ** 2), r2)
        fr[m] = np.clip(200 - depth * (1 - 0.9 * np.sqrt(core[m]).clip(0, 1)) - rng.integers(0, 4, int(m.sum())), 0, 199).astype(np.uint8)
    fr[5, 5:8] = 150; fr[150, 200] = 120                   # tiny blobs: no midline at any threshold
    return fr, bg


@pytest.mark.parametrize("T0,resample", [(12, 1.0), (30, 1.0), (12, 0.6)])
def test_threshold_loop_against_the_compiled_reference(ref, T0, resample):
    P = posture.default_params()
    set_ref_settings(ref, P)
    ref.ref_posture_settings(int(T0), C.c_float(resample))
    ref.ref_background_settings(1, 1, 0)                    # track_threshold_is_absolute, track_background_subtraction, gray
    n_mid = n_outline = n_none = n_retry = 0
    for seed in (4, 5, 6):
        fr, bg = graded_frame(seed)
        blobs = seg.segment_frame(fr, bg, seg.Params(detect_threshold=10, detect_size_filter=[]))
        for b in range(len(blobs)):
            l, p = blobs.blob(b)
            p = np.asarray(p)
            raw = np.zeros((len(l), 4), np.uint16); raw[:, 0], raw[:, 1], raw[:, 2] = l["x0"], l["x1"], l["y"]
            cap = 4 * len(p) + 64
            pts = np.zeros((cap, 2), np.float32); segs = np.zeros((cap, 4), np.float32)
            n_pts, tail, head = C.c_int64(), C.c_int64(-1), C.c_int64(-1)
            k = ref.ref_calculate_posture(_p(raw), C.c_int64(len(raw)), _p(p), C.c_int64(len(p)), 1, _p(bg), bg.shape[1], bg.shape[0], 1, 0,
                                          _p(pts), C.c_int64(cap), C.byref(n_pts), _p(segs), C.c_int64(cap), C.byref(tail), C.byref(head))
            try:
                mine = posture.calculate_posture(l, p, bg, track_posture_threshold=T0, outline_resample=resample, method=seg.DIFF_ABSOLUTE, params=P)
            except ValueError:
                assert k == -2, (seed, b, k)
                n_none += 1
                continue
            assert k != -2, (seed, b)
            want_outline = pts[:n_pts.value]
            assert mine["outline"].shape == want_outline.shape and np.array_equal(mine["outline"].view(np.uint32), want_outline.view(np.uint32)), (seed, b)
            if mine["segments"] is None:
                assert k == -1, (seed, b, k)
                n_outline += 1
            else:
                assert k == len(mine["segments"]) and (tail.value, head.value) == (mine["tail"], mine["head"]), (seed, b, k)
                assert np.array_equal(mine["segments"].view(np.uint32), segs[:k].view(np.uint32)), (seed, b)
                n_mid += 1
                n_retry += int(mine["threshold"] > T0)
    print(f"T0={T0} resample={resample}: {n_mid} midlines ({n_retry} after retries), {n_outline} outline-only, {n_none} without posture")
    assert n_mid > 25 and n_none >= 1


def test_random_posture_settings_against_the_compiled_reference(ref):
    """25 seeded random settings (curvature range, pointy / broad tails, walk offset, 0 ... 8 Fourier harmonics, smoothing window and step, midline_start_with_head,
    midline_invert, stiffness, 5 ... 40 midline points, posture threshold 5 ... 60, outline_resample 0.5 ... 2) x every blob of a graded frame: outcome, outline,
    segments, tail / head of posture::calculate_posture -- the compiled reference and the oracle agree bit for bit."""
    rng = np.random.default_rng(7)
    total = 0
    for _ in range(25):
        P = posture.default_params()
        P.outline_curvature_range_ratio = float(rng.choice([0.01, 0.03, 0.05, 0.1]))
        P.peak_mode = int(rng.integers(0, 2)); P.midline_walk_offset = float(rng.choice([0.0, 0.025, 0.05, 0.1]))
        P.outline_approximate = int(rng.choice([0, 1, 3, 5, 8])); P.outline_smooth_samples = int(rng.choice([0, 2, 4, 8])); P.outline_smooth_step = int(rng.choice([1, 1, 2]))
        P.midline_start_with_head = int(rng.integers(0, 2)); P.midline_invert = int(rng.integers(0, 2))
        P.midline_stiff_percentage = float(rng.choice([0.0, 0.15, 0.4])); P.midline_resolution = int(rng.choice([5, 12, 25, 40]))
        T0 = int(rng.choice([5, 12, 30, 60])); resample = float(rng.choice([0.5, 1.0, 1.0, 2.0]))
        set_ref_settings(ref, P)
        ref.ref_posture_settings(T0, C.c_float(resample))
        ref.ref_background_settings(1, 1, 0)
        fr, bg = graded_frame(int(rng.integers(0, 100)))
        blobs = seg.segment_frame(fr, bg, seg.Params(detect_threshold=10, detect_size_filter=[]))
        for b in range(len(blobs)):
            l, p = blobs.blob(b)
            p = np.asarray(p)
            raw = np.zeros((len(l), 4), np.uint16); raw[:, 0], raw[:, 1], raw[:, 2] = l["x0"], l["x1"], l["y"]
            cap = 4 * len(p) + 64
            pts = np.zeros((cap, 2), np.float32); segs = np.zeros((cap, 4), np.float32)
            n_pts, tail, head = C.c_int64(), C.c_int64(-1), C.c_int64(-1)
            k = ref.ref_calculate_posture(_p(raw), C.c_int64(len(raw)), _p(p), C.c_int64(len(p)), 1, _p(bg), bg.shape[1], bg.shape[0], 1, 0,
                                          _p(pts), C.c_int64(cap), C.byref(n_pts), _p(segs), C.c_int64(cap), C.byref(tail), C.byref(head))
            total += 1
            try:
                mine = posture.calculate_posture(l, p, bg, track_posture_threshold=T0, outline_resample=resample, method=seg.DIFF_ABSOLUTE, params=P)
            except ValueError:
                assert k == -2, (b, k)
                continue
            assert k != -2, b
            assert mine["outline"].shape == pts[:n_pts.value].shape and np.array_equal(mine["outline"].view(np.uint32), pts[:n_pts.value].view(np.uint32)), b
            if mine["segments"] is None:
                assert k == -1, (b, k)
            else:
                assert k == len(mine["segments"]) and (tail.value, head.value) == (mine["tail"], mine["head"]), (b, k)
                assert np.array_equal(mine["segments"].view(np.uint32), segs[:k].view(np.uint32)), b
    set_ref_settings(ref, posture.default_params())
    assert total > 300
