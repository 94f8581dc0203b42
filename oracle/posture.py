"""From a resampled outline to the raw, post-processed and normalised midline and the posture-normalised crop -- TEST INFRASTRUCTURE.

Python face of the restatement in trex_oracle.c of
  Outline::smooth / offset_to_middle / calculate_midline   Application/src/tracker/tracking/Outline.cpp:330-452,454-718,768-868
  Midline::post_process / normalize / fix_length / transform  Application/src/tracker/tracking/Outline.cpp:870-1085,1113-1456
  image::normalize_image (posture / legacy)                Application/src/tracker/tracking/FilterCache.cpp:21-115,133-154,266-276
  periodic::curvature / differentiate / find_peaks / eft / ieft, fast::cos
                                                            Application/src/commons/common/misc/CircularGraph.cpp:12-606
Pinned on the reference's own code: tests/test_oracle_ref_circular_graph.py and tests/test_oracle_ref_outline.py compare these functions with
CircularGraph.cpp / Outline.cpp / gui/Transform.cpp compiled unmodified from the reference checkout (oracle/build_ref.py) bit for bit;
tests/test_oracle_posture.py additionally checks each piece against an independent numpy formulation and the whole chain on shapes whose midline
is known by construction.  pixel::find_outer_points (oracle/seg.py) is pinned the same way (tests/test_oracle_ref_pixeltree.py); not covered by
that build: the threshold loop of Posture.cpp."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .seg import _p, lib


class PostureParams(C.Structure):
    _fields_ = [("outline_smooth_samples", C.c_int32), ("outline_smooth_step", C.c_int32), ("outline_approximate", C.c_int32),
                ("outline_curvature_range_ratio", C.c_float), ("midline_walk_offset", C.c_float), ("peak_mode", C.c_int32),
                ("midline_start_with_head", C.c_int32), ("midline_invert", C.c_int32),
                ("midline_resolution", C.c_int32), ("midline_stiff_percentage", C.c_float)]


PEAK_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("width", "<f4"), ("integral", "<f4"), ("r0", "<f4"), ("r1", "<f4"),
                       ("max_y_extrema", "<f4"), ("max_y", "<f4"), ("n_pts", "<i4")])


def _lib():
    L = lib()
    if not getattr(L, "_posture_ready", False):
        vp = C.c_void_p
        L.to_posture_default_params.argtypes = [C.POINTER(PostureParams)]; L.to_posture_default_params.restype = None
        L.to_fast_cos.argtypes = [C.c_float]; L.to_fast_cos.restype = C.c_float
        L.to_outline_smooth.argtypes = [vp, C.c_int64, C.c_int, C.c_int, vp]; L.to_outline_smooth.restype = C.c_int
        L.to_periodic_curvature.argtypes = [vp, C.c_int64, C.c_int, C.c_int, vp]; L.to_periodic_curvature.restype = None
        L.to_orientation_sum.argtypes = [vp, C.c_int64]; L.to_orientation_sum.restype = C.c_float
        L.to_eft.argtypes = [vp, C.c_int64, C.c_int, vp]; L.to_eft.restype = None
        L.to_ieft.argtypes = [vp, C.c_int, C.c_int64, C.c_float, C.c_float, vp]; L.to_ieft.restype = None
        L.to_find_peaks.argtypes = [vp, C.c_int64, C.c_int, vp, C.c_int64]; L.to_find_peaks.restype = C.c_int64
        L.to_offset_to_middle.argtypes = [vp, C.c_int64, C.POINTER(PostureParams), C.POINTER(C.c_int64), C.POINTER(C.c_int64), vp]
        L.to_offset_to_middle.restype = C.c_int
        L.to_calculate_midline.argtypes = [vp, C.c_int64, C.POINTER(PostureParams), vp, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.to_calculate_midline.restype = C.c_int64
        L.to_midline_post_process.argtypes = [vp, C.c_int64, C.POINTER(PostureParams), vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.to_midline_post_process.restype = C.c_int
        L.to_midline_normalize.argtypes = [vp, C.c_int64, C.POINTER(PostureParams), C.c_float, vp, vp]; L.to_midline_normalize.restype = C.c_int64
        L.to_posture_matrix.argtypes = [C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, vp]; L.to_posture_matrix.restype = None
        L.to_crop_blob_posture.argtypes = [vp, C.c_int64, vp, vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                                           C.c_int, C.c_int, vp]
        L.to_crop_blob_posture.restype = C.c_int
        L._posture_ready = True
    return L


def default_params(**kw) -> PostureParams:
    p = PostureParams()
    _lib().to_posture_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def fast_cos(x: float) -> float:
    return float(_lib().to_fast_cos(C.c_float(x)))


def smooth(points, samples=4, step=1):
    pts = np.ascontiguousarray(points, np.float32)
    out = np.empty_like(pts)
    return out if _lib().to_outline_smooth(_p(pts), len(pts), samples, step, _p(out)) else pts.copy()


def curvature(points, r, absolute=False):
    pts = np.ascontiguousarray(points, np.float32)
    out = np.zeros(len(pts), np.float32)
    _lib().to_periodic_curvature(_p(pts), len(pts), int(r), int(absolute), _p(out))
    return out


def orientation_sum(points) -> float:
    pts = np.ascontiguousarray(points, np.float32)
    return float(_lib().to_orientation_sum(_p(pts), len(pts)))


def eft(points, order=3):
    pts = np.ascontiguousarray(points, np.float32)
    out = np.zeros((order, 4), np.float32)
    _lib().to_eft(_p(pts), len(pts), order, _p(out))
    return out


def ieft(coeffs, n_points, offset=(0.0, 0.0)):
    c = np.ascontiguousarray(coeffs, np.float32)
    out = np.zeros((n_points, 2), np.float32)
    _lib().to_ieft(_p(c), len(c), n_points, C.c_float(offset[0]), C.c_float(offset[1]), _p(out))
    return out


def find_peaks(values, broad=False):
    v = np.ascontiguousarray(values, np.float32)
    out = np.zeros(len(v) + 1, PEAK_DTYPE)
    n = _lib().to_find_peaks(_p(v), len(v), int(broad), _p(out), len(out))
    if n < 0:
        raise MemoryError
    return out[:n].copy()


def offset_to_middle(points, params: PostureParams | None = None):
    """Returns (points after reversal / approximation / rotation, tail index, head index, curvature)."""
    pts = np.ascontiguousarray(points, np.float32).copy()
    P = params or default_params()
    t, h = C.c_int64(), C.c_int64()
    curv = np.zeros(len(pts), np.float32)
    rc = _lib().to_offset_to_middle(_p(pts), len(pts), C.byref(P), C.byref(t), C.byref(h), _p(curv))
    if rc:
        raise ValueError(f"offset_to_middle failed ({rc})")
    return pts, int(t.value), int(h.value), curv


class MidlineError(ValueError):
    """std::unexpected of Outline::calculate_midline; .points = the outline as the failed call left it (smoothed, approximated, rotated: the
    reference works on the Outline in place, Outline.cpp:768-777)."""
    def __init__(self, msg, points):
        super().__init__(msg)
        self.points = points


def calculate_midline(points, params: PostureParams | None = None):
    """Returns (segments (n,4): pos.x, pos.y, height, l_length; tail index; head index; the outline as the walk saw it),
    or raises MidlineError (a ValueError) like the reference returns std::unexpected."""
    pts = np.ascontiguousarray(points, np.float32).copy()
    P = params or default_params()
    seg = np.zeros((len(pts) + 4, 4), np.float32)
    t, h = C.c_int64(), C.c_int64()
    n = _lib().to_calculate_midline(_p(pts), len(pts), C.byref(P), _p(seg), len(seg), C.byref(t), C.byref(h))
    if n < 0:
        raise MidlineError({-1: "Empty outline was given, cannot calculate midline.", -2: "Too few midline segments calculated.",
                            -4: "capacity"}[int(n)], pts)
    return seg[:n].copy(), int(t.value), int(h.value), pts


def calculate_posture(lines, pixels, bg, track_posture_threshold=0, outline_resample=1.0, method=None, params: PostureParams | None = None):
    """posture::calculate_posture(Frame_t, pv::BlobWeakPtr) (T/tracking/Posture.cpp:305-400) with posture_closing_steps = 0: starting at
    track_posture_threshold, threshold the blob (pixel::threshold_get_biggest_blob, C/processing/PixelTree.cpp:297-340: the
    sub-blob with the most pixels, the first one among equals -- here in the oracle's canonical blob order), take its longest
    outline in the frame of the ORIGINAL blob's bounds (:337), resample it and try calculate_midline; on failure raise the
    threshold by 2 until the sub-blob has fewer than max(1, pixels / 10) pixels or the threshold reaches the start + 100.
    Returns dict(outline=(n,2) points, segments=(m,4) or None, tail, head, threshold): like the reference, a posture without a
    midline still carries the first resampled outline (:386-397); raises ValueError("Cannot find valid posture.") otherwise."""
    from . import seg
    method = seg.DIFF_ABSOLUTE if method is None else method
    lines = np.ascontiguousarray(lines); pixels = np.ascontiguousarray(pixels, np.uint8)
    npx = int((lines["x1"].astype(np.int64) - lines["x0"] + 1).sum())
    minimum_pixels = max(1, npx // 10)
    ox, oy = int(lines["x0"].min()), int(lines["y"].min())
    one = seg.Blobs(lines, pixels, np.array([0, len(lines)], np.int64), np.array([0, len(pixels)], np.int64))
    threshold, first_outline = int(track_posture_threshold), None
    while True:
        sub = seg.rethreshold(one, bg, threshold, method, keep_single=True)      # threshold_get_biggest_blob sees every sub-blob
        sizes = np.diff(sub.px_off)
        n_sub = 0
        if len(sub):
            k = int(np.argmax(sizes))                       # first maximum
            sl, _ = sub.blob(k)
            n_sub = int(sizes[k])
            raw = seg.longest_outline(sl)
            if len(raw):
                raw = raw + np.array([int(sl["x0"].min()) - ox, int(sl["y"].min()) - oy], np.float32)
                pts = seg.outline_resample(raw, outline_resample)
                try:
                    segs, tail, head, walked = calculate_midline(pts, params)
                    return dict(outline=walked, segments=segs, tail=tail, head=head, threshold=threshold)
                except MidlineError as e:                       # the failed call has already smoothed / rotated the outline in place (:358-366 keeps THOSE points)
                    if first_outline is None and len(e.points):
                        first_outline = e.points
        threshold += 2
        if n_sub < minimum_pixels or threshold >= track_posture_threshold + 100:
            break
    if first_outline is not None:
        return dict(outline=first_outline, segments=None, tail=-1, head=-1, threshold=None)
    raise ValueError("Cannot find valid posture.")


def post_process(segments, params: PostureParams | None = None, move_dir=None, tail=-1, head=-1):
    """Midline::post_process (Outline.cpp:895-1062) on raw midline segments (n,4).  Returns (segments, tail, head, inverted_because_previous);
    raises IndexError where the reference's segments().at() throws."""
    seg = np.ascontiguousarray(segments, np.float32).copy()
    P = params or default_params()
    t, h = C.c_int64(tail), C.c_int64(head)
    md = None if move_dir is None else np.ascontiguousarray(move_dir, np.float32)
    rc = _lib().to_midline_post_process(_p(seg), len(seg), C.byref(P), _p(md) if md is not None else None, C.byref(t), C.byref(h))
    if rc < 0:
        raise IndexError("segments().at(i + 1) out of range in Midline::post_process")
    return seg, int(t.value), int(h.value), bool(rc)


def normalize(segments, params: PostureParams | None = None, fix_length=-1.0):
    """Midline::normalize (Outline.cpp:1268-1456).  Returns (segments (resolution,4), len, angle, offset (2,)) or None (nullptr)."""
    seg = np.ascontiguousarray(segments, np.float32)
    P = params or default_params()
    out = np.zeros((int(P.midline_resolution) + 2, 4), np.float32)
    info = np.zeros(4, np.float32)
    n = _lib().to_midline_normalize(_p(seg), len(seg), C.byref(P), C.c_float(fix_length), _p(out), _p(info))
    if n <= 0:
        return None
    return out[:n].copy(), float(info[0]), float(info[1]), info[2:4].copy()


def posture_matrix(angle, offset, midline_length, out_size=(80, 80), image_scale=1.0, legacy=False):
    M = np.zeros(6, np.float64)
    _lib().to_posture_matrix(C.c_float(angle), C.c_float(offset[0]), C.c_float(offset[1]), C.c_float(midline_length), C.c_float(image_scale),
                             int(legacy), int(out_size[0]), int(out_size[1]), _p(M))
    return M.reshape(2, 3)


def crop_blob_posture(lines, pixels, bg, method, angle, offset, midline_length, out_size=(80, 80), image_scale=1.0, legacy=False):
    """constraints::diff_image(posture | legacy) (FilterCache.cpp:266-276): the blob's (difference) image warped by the midline transform."""
    lines = np.ascontiguousarray(lines); pixels = np.ascontiguousarray(pixels, np.uint8); bg = np.ascontiguousarray(bg, np.uint8)
    out = np.zeros((out_size[1], out_size[0]), np.uint8)
    ok = _lib().to_crop_blob_posture(_p(lines), len(lines), _p(pixels), _p(bg), bg.shape[1], int(method), C.c_float(angle), C.c_float(offset[0]),
                                     C.c_float(offset[1]), C.c_float(midline_length), C.c_float(image_scale), int(legacy), int(out_size[0]),
                                     int(out_size[1]), _p(out))
    return out if ok else None
