"""ctypes front-end of liboracle.so (oracle/trex_oracle.c).  TEST INFRASTRUCTURE ONLY.

Restates, on the CPU, what the reference does in
  RawProcessing::generate_binary   Application/src/commons/common/processing/RawProcessing.cpp:263-600
  Source::extract_lines            .../processing/Source.cpp:156-255
  merge_lines / run_fast           .../processing/CPULabeling.cpp:44-343
  BackgroundSubtraction::apply     Application/src/tracker/python/BackgroundSubtraction.cpp:209-313
  image::calculate_diff_image      Application/src/tracker/tracking/FilterCache.cpp:158-235
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ORDER_CANONICAL, ORDER_REF_LAZY, ORDER_REF_ABSORB = 0, 1, 2
DIFF_NONE, DIFF_ABSOLUTE, DIFF_SIGN = 0, 1, 2

LINE_DTYPE = np.dtype([("x0", "<u2"), ("x1", "<u2"), ("y", "<u2"), ("pad", "<u2")])


class _Params(C.Structure):
    _fields_ = [
        ("detect_threshold", C.c_int32), ("threshold_maximum", C.c_int32),
        ("enable_difference", C.c_int32), ("detect_threshold_is_absolute", C.c_int32),
        ("image_invert", C.c_int32), ("use_closing", C.c_int32), ("closing_size", C.c_int32),
        ("dilation_size", C.c_int32), ("cm_per_pixel", C.c_float), ("n_size_ranges", C.c_int32),
        ("size_lo", C.c_double * 4), ("size_hi", C.c_double * 4),
        ("blur_difference", C.c_int32), ("use_adaptive_threshold", C.c_int32), ("adaptive_threshold_scale", C.c_float), ("open_size", C.c_int32),
    ]


@dataclass
class Params:
    """The settings keys the path reads, with the reference defaults."""
    detect_threshold: int = 15
    threshold_maximum: int = 255
    enable_difference: bool = True
    detect_threshold_is_absolute: bool = True
    image_invert: bool = False
    use_closing: bool = False
    closing_size: int = 3
    dilation_size: int = 0
    cm_per_pixel: float = 1.0
    detect_size_filter: list = field(default_factory=lambda: [(10.0, 100000.0)])
    blur_difference: bool = False
    use_adaptive_threshold: bool = False
    adaptive_threshold_scale: float = 2.0
    open_size: int = 0          # not a reference setting: north_star's optional n x n open of the threshold mask (0 = off)

    def c(self) -> _Params:
        p = _Params()
        p.detect_threshold = self.detect_threshold
        p.threshold_maximum = self.threshold_maximum
        p.enable_difference = int(self.enable_difference)
        p.detect_threshold_is_absolute = int(self.detect_threshold_is_absolute)
        p.image_invert = int(self.image_invert)
        p.use_closing = int(self.use_closing)
        p.closing_size = self.closing_size
        p.dilation_size = self.dilation_size
        p.cm_per_pixel = self.cm_per_pixel
        p.blur_difference = int(self.blur_difference)
        p.use_adaptive_threshold = int(self.use_adaptive_threshold)
        p.adaptive_threshold_scale = self.adaptive_threshold_scale
        p.open_size = int(self.open_size)
        p.n_size_ranges = len(self.detect_size_filter)
        for i, (lo, hi) in enumerate(self.detect_size_filter):
            p.size_lo[i], p.size_hi[i] = lo, hi
        return p


def build(force: bool = False) -> str:
    """Compile liboracle.so (and oracle/_ref when /root/reference exists) with the Makefile."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "trex_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u8p, i64p, i32p, vp = C.POINTER(C.c_uint8), C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.c_void_p
        L.to_generate_binary.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(_Params), vp]
        L.to_generate_binary.restype = C.c_int
        L.to_extract_lines.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int64]
        L.to_extract_lines.restype = C.c_int64
        L.to_label_runs.argtypes = [vp, C.c_int64, C.c_int, vp]
        L.to_label_runs.restype = C.c_int64
        L.to_segment_frame.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(_Params), C.c_int,
                                       vp, C.c_int64, vp, C.c_int64, vp, vp, C.c_int64, vp]
        L.to_segment_frame.restype = C.c_int64
        L.to_blob_id.argtypes = [vp, C.c_int64]
        L.to_blob_id.restype = C.c_uint32
        L.to_image_from_lines.argtypes = [vp, C.c_int64, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
        L.to_image_from_lines.restype = C.c_int64
        L.to_crop_blob.argtypes = [vp, C.c_int64, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]
        L.to_crop_blob.restype = None
        L.to_segment_batch.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.POINTER(_Params), C.c_int,
                                       C.c_int, C.c_int, C.c_int, vp, vp, C.c_int]
        L.to_segment_batch.restype = C.c_int64
        L.to_num_threads.restype = C.c_int
        L.to_rethreshold_frame.argtypes = [vp, vp, vp, vp, C.c_int64, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int64, vp, C.c_int64, vp, vp, C.c_int64]
        L.to_rethreshold_frame.restype = C.c_int64
        L.to_blob_orientation.argtypes = [vp, C.c_int64, vp]
        L.to_blob_orientation.restype = C.c_float
        L.to_moments_matrix.argtypes = [C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, vp]
        L.to_moments_matrix.restype = None
        L.to_warp_affine_u8.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int]
        L.to_warp_affine_u8.restype = None
        L.to_crop_blob_moments.argtypes = [vp, C.c_int64, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]
        L.to_crop_blob_moments.restype = None
        L.to_rethreshold_frame_rgb.argtypes = L.to_rethreshold_frame.argtypes
        L.to_rethreshold_frame_rgb.restype = C.c_int64
        L.to_average.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]
        L.to_average.restype = C.c_int
        L.to_find_outer_points.argtypes = [vp, C.c_int64, vp, C.c_int64, vp, C.c_int64]; L.to_find_outer_points.restype = C.c_int64
        L.to_outline_resample.argtypes = [vp, C.c_int64, C.c_float, vp, C.c_int64]; L.to_outline_resample.restype = C.c_int64
        L.to_box_mean.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]; L.to_box_mean.restype = C.c_int
        L.to_adaptive_neighbourhood.argtypes = [C.c_int, C.c_float]; L.to_adaptive_neighbourhood.restype = C.c_int
        L.to_bgr2gray.argtypes = [vp, C.c_int64, C.c_int, vp]
        L.to_bgr2gray.restype = None
        L.to_bgr2gray_tracker.argtypes = [vp, C.c_int64, vp]
        L.to_convert_to_r3g3b2.argtypes = [vp, C.c_int64, C.c_int, vp]; L.to_convert_to_r3g3b2.restype = None
        L.to_convert_from_r3g3b2.argtypes = [vp, C.c_int64, vp]; L.to_convert_from_r3g3b2.restype = None
        L.to_bgr2gray_tracker.restype = None
        L.to_generate_binary_color.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.POINTER(_Params), vp, vp]
        L.to_generate_binary_color.restype = C.c_int
        L.to_segment_frame_color.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.POINTER(_Params), C.c_int,
                                             vp, C.c_int64, vp, C.c_int64, vp, vp, C.c_int64, vp]
        L.to_segment_frame_color.restype = C.c_int64
        L.to_image_from_lines_rgb.argtypes = [vp, C.c_int64, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
        L.to_image_from_lines_rgb.restype = C.c_int64
        L.to_crop_blob_rgb.argtypes = [vp, C.c_int64, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]
        L.to_crop_blob_rgb.restype = None
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def box_mean(img: np.ndarray, k: int, border: str = "replicate") -> np.ndarray:
    """cv::boxFilter(normalize=true) / cv::blur of an 8-bit image; border "replicate" or "reflect101"."""
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    if lib().to_box_mean(_p(img), img.shape[1], img.shape[0], int(k), {"replicate": 0, "reflect101": 1}[border], _p(out)) != 0:
        raise MemoryError
    return out


def adaptive_neighbourhood(cols: int, scale: float) -> int:
    return int(lib().to_adaptive_neighbourhood(int(cols), C.c_float(scale)))


def generate_binary(frame: np.ndarray, bg: np.ndarray, params: Params) -> np.ndarray:
    frame = np.ascontiguousarray(frame, np.uint8); bg = np.ascontiguousarray(bg, np.uint8)
    h, w = frame.shape
    out = np.empty_like(frame)
    pc = params.c()
    if lib().to_generate_binary(_p(frame), _p(bg), w, h, C.byref(pc), _p(out)) != 0:
        raise MemoryError
    return out


@dataclass
class Blobs:
    """SoA blob list of one frame (same content as the reference's blobs_t / blob::Pair list)."""
    lines: np.ndarray      # LINE_DTYPE [L]
    pixels: np.ndarray     # u8 [P]
    line_off: np.ndarray   # i64 [K+1]
    px_off: np.ndarray     # i64 [K+1]

    def __len__(self):
        return len(self.line_off) - 1

    def blob(self, k):
        return (self.lines[self.line_off[k]:self.line_off[k + 1]],
                self.pixels[self.px_off[k]:self.px_off[k + 1]])

    def as_set(self):
        """Order-independent identity: {(lines bytes, pixel bytes)} (set equality)."""
        return {(self.lines[self.line_off[k]:self.line_off[k + 1]].tobytes(),
                 self.pixels[self.px_off[k]:self.px_off[k + 1]].tobytes()) for k in range(len(self))}

    def as_list(self):
        return [(self.lines[self.line_off[k]:self.line_off[k + 1]].tobytes(),
                 self.pixels[self.px_off[k]:self.px_off[k + 1]].tobytes()) for k in range(len(self))]


def segment_frame(frame, bg, params: Params, order=ORDER_CANONICAL, want_binary=False):
    frame = np.ascontiguousarray(frame, np.uint8); bg = np.ascontiguousarray(bg, np.uint8)
    h, w = frame.shape
    capL, capP, capB = 1 << 16, max(1 << 16, w * h // 8), 1 << 14
    binary = np.empty_like(frame) if want_binary else None
    pc = params.c()
    while True:
        lines = np.zeros(capL, LINE_DTYPE); px = np.zeros(capP, np.uint8)
        lo = np.zeros(capB + 1, np.int64); po = np.zeros(capB + 1, np.int64)
        k = lib().to_segment_frame(_p(frame), _p(bg), w, h, C.byref(pc), order, _p(lines), capL,
                                   _p(px), capP, _p(lo), _p(po), capB, _p(binary))
        if k >= 0:
            break
        if k == -1:
            raise MemoryError
        need = -k
        capL = max(capL, need); capP = max(capP, need); capB = max(capB, need)
    b = Blobs(lines[:lo[k]].copy(), px[:po[k]].copy(), lo[:k + 1].copy(), po[:k + 1].copy())
    return (b, binary) if want_binary else b


def label_image(img, order=ORDER_CANONICAL):
    """CPULabeling::run on a binary/grey image: no threshold, no size filter."""
    p = Params(detect_threshold=0, enable_difference=False, detect_size_filter=[])
    return segment_frame(img, img, p, order)


def blob_id(lines: np.ndarray) -> int:
    lines = np.ascontiguousarray(lines)
    return int(lib().to_blob_id(_p(lines), len(lines)))


def image_from_lines(lines, pixels, bg, method=DIFF_NONE, base_threshold=0):
    lines = np.ascontiguousarray(lines); pixels = np.ascontiguousarray(pixels, np.uint8)
    bg = np.ascontiguousarray(bg, np.uint8)
    w = int(lines["x1"].max()) - int(lines["x0"].min()) + 1
    h = int(lines["y"].max()) - int(lines["y"].min()) + 1
    rect = np.zeros(4, np.int32)
    mask = np.zeros((h, w), np.uint8); grey = np.zeros((h, w), np.uint8); diff = np.zeros((h, w), np.uint8)
    n = lib().to_image_from_lines(_p(lines), len(lines), _p(pixels), _p(bg), bg.shape[1], method,
                                  base_threshold, _p(rect), _p(mask), _p(grey), _p(diff))
    return rect, int(n), mask, grey, diff


def crop_blob(lines, pixels, bg, method=DIFF_ABSOLUTE, out_w=80, out_h=80):
    lines = np.ascontiguousarray(lines); pixels = np.ascontiguousarray(pixels, np.uint8)
    bg = np.ascontiguousarray(bg, np.uint8)
    out = np.zeros((out_h, out_w), np.uint8)
    lib().to_crop_blob(_p(lines), len(lines), _p(pixels), _p(bg), bg.shape[1], method, out_w, out_h, _p(out))
    return out


ENC_GRAY, ENC_RGB8, ENC_R3G3B2 = 0, 1, 2


def resize_nearest(img: np.ndarray, factor: float) -> np.ndarray:
    """cv::resize(src, dst, Size(), f, f, INTER_NEAREST) -- resize_image's default (C/misc/detail.h:465-469): dsize = cvRound(n * f),
    source index = min(floor(d * (1 / f)), n - 1).  Checked against cv2 in tests/test_oracle_golden.py."""
    f = float(np.float32(factor))                       # individual_image_scale is a float setting, widened to double
    h, w = img.shape[:2]
    dw, dh = int(np.rint(w * f)), int(np.rint(h * f))
    ifx = 1.0 / f
    xs = np.minimum(np.floor(np.arange(dw) * ifx).astype(np.int64), w - 1)
    ys = np.minimum(np.floor(np.arange(dh) * ifx).astype(np.int64), h - 1)
    return img[ys][:, xs]


def crop_blob_scaled(lines, pixels, bg, method=DIFF_ABSOLUTE, scale=1.0, out_w=80, out_h=80):
    """calculate_diff_image with individual_image_scale != 1 (T/tracking/FilterCache.cpp:176-228): the masked blob image is
    resized (resize_image, INTER_NEAREST) before the centre pad / centre crop to the output size."""
    _, _, mask, img, diff = image_from_lines(lines, pixels, bg, method, 0)
    src = img if method == DIFF_NONE else diff          # image.copyTo(padded, mask): both are zero outside the mask
    out = np.zeros((out_h, out_w), np.uint8)
    if float(np.float32(scale)) != 1.0:
        src = resize_nearest(src, scale)
    h, w = src.shape
    if h == 0 or w == 0:
        return out
    def place(n, m):        # (offset in the output, offset in the source, count) along one axis: pad left = d - d/2, crop start = d - d/2
        if n < m:
            d = m - n; return d - d // 2, 0, n
        d = n - m; return 0, d - d // 2, m
    ox, sx, cw = place(w, out_w); oy, sy, ch = place(h, out_h)
    out[oy:oy + ch, ox:ox + cw] = src[sy:sy + ch, sx:sx + cw]
    return out


def bgr2gray(img: np.ndarray) -> np.ndarray:
    """cv::cvtColor(BGR2GRAY / BGRA2GRAY) on u8 (H,W,3|4) -> (H,W)."""
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty(img.shape[:-1], np.uint8)
    lib().to_bgr2gray(_p(img), out.size, img.shape[-1], _p(out))
    return out


def bgr2gray_tracker(img: np.ndarray) -> np.ndarray:
    """cmn::bgr2gray (Background.h:76-81) on u8 (...,3)."""
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty(img.shape[:-1], np.uint8)
    lib().to_bgr2gray_tracker(_p(img), out.size, _p(out))
    return out


def convert_to_r3g3b2(img: np.ndarray) -> np.ndarray:
    """convert_to_r3g3b2<3|4> (C/misc/detail.h:533-555): (H,W,3|4) B,G,R(,A) -> (H,W) codes."""
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty(img.shape[:-1], np.uint8)
    lib().to_convert_to_r3g3b2(_p(img), out.size, img.shape[-1], _p(out))
    return out


def convert_from_r3g3b2(codes: np.ndarray) -> np.ndarray:
    """convert_from_r3g3b2<3> / r3g3b2_to_vec (C/misc/detail.h:515-531): codes (...) -> (...,3)."""
    codes = np.ascontiguousarray(codes, np.uint8)
    out = np.empty(codes.shape + (3,), np.uint8)
    lib().to_convert_from_r3g3b2(_p(codes), codes.size, _p(out))
    return out


def crop_blob_r3g3b2(lines, pixels, bg_codes, method=DIFF_ABSOLUTE, out_w=80, out_h=80):
    """calculate_diff_image for an r3g3b2 blob: imageFromLines renders 3 channels (Background.cpp:134-139), each pixel code
    expanded by r3g3b2_to_vec (Background.h:113-116) and differenced per channel against the tracker's background, which
    Background's constructor expanded the same way (Background.cpp:65-69)."""
    return crop_blob_rgb(lines, convert_from_r3g3b2(np.ascontiguousarray(pixels, np.uint8)).reshape(-1),
                         convert_from_r3g3b2(bg_codes), method, out_w, out_h)


def generate_binary_color(frame, bg, params: Params, encoding=ENC_GRAY, color_channel=-1):
    """frame (H,W,C) u8, bg (H,W) for gray / r3g3b2 encoding, (H,W,3) for rgb8 -> (H,W) or (H,W,3), grey plane."""
    frame = np.ascontiguousarray(frame, np.uint8); bg = np.ascontiguousarray(bg, np.uint8)
    h, w, cn = frame.shape
    out = np.empty((h, w, 3) if encoding == ENC_RGB8 else (h, w), np.uint8)
    gray = np.empty((h, w), np.uint8)
    pc = params.c()
    if lib().to_generate_binary_color(_p(frame), cn, encoding, color_channel, _p(bg), w, h, C.byref(pc), _p(out), _p(gray)) != 0:
        raise ValueError("generate_binary_color failed (channels / encoding)")
    return out, gray


def segment_frame_color(frame, bg, params: Params, encoding=ENC_GRAY, color_channel=-1, order=ORDER_CANONICAL):
    """BackgroundSubtraction::apply's body for one colour frame; Blobs.pixels holds 1 (gray) or 3 (rgb8) bytes per pixel."""
    frame = np.ascontiguousarray(frame, np.uint8); bg = np.ascontiguousarray(bg, np.uint8)
    h, w, cn = frame.shape
    capL, capP, capB = 1 << 16, max(1 << 16, w * h // 4), 1 << 14
    pc = params.c()
    while True:
        lines = np.zeros(capL, LINE_DTYPE); px = np.zeros(capP, np.uint8)
        lo = np.zeros(capB + 1, np.int64); po = np.zeros(capB + 1, np.int64)
        k = lib().to_segment_frame_color(_p(frame), cn, encoding, color_channel, _p(bg), w, h, C.byref(pc), order,
                                         _p(lines), capL, _p(px), capP, _p(lo), _p(po), capB, None)
        if k >= 0:
            break
        if k == -1:
            raise MemoryError
        need = -k * 3
        capL = max(capL, need); capP = max(capP, need); capB = max(capB, need)
    return Blobs(lines[:lo[k]].copy(), px[:po[k]].copy(), lo[:k + 1].copy(), po[:k + 1].copy())


def image_from_lines_rgb(lines, pixels, bg3, method=DIFF_NONE, base_threshold=0):
    lines = np.ascontiguousarray(lines); pixels = np.ascontiguousarray(pixels, np.uint8)
    bg3 = np.ascontiguousarray(bg3, np.uint8)
    w = int(lines["x1"].max()) - int(lines["x0"].min()) + 1
    h = int(lines["y"].max()) - int(lines["y"].min()) + 1
    rect = np.zeros(4, np.int32)
    mask = np.zeros((h, w), np.uint8); img = np.zeros((h, w, 3), np.uint8); diff = np.zeros((h, w, 3), np.uint8)
    n = lib().to_image_from_lines_rgb(_p(lines), len(lines), _p(pixels), _p(bg3), bg3.shape[1], method,
                                      base_threshold, _p(rect), _p(mask), _p(img), _p(diff))
    return rect, int(n), mask, img, diff


def crop_blob_rgb(lines, pixels, bg3, method=DIFF_ABSOLUTE, out_w=80, out_h=80):
    lines = np.ascontiguousarray(lines); pixels = np.ascontiguousarray(pixels, np.uint8)
    bg3 = np.ascontiguousarray(bg3, np.uint8)
    out = np.zeros((out_h, out_w, 3), np.uint8)
    lib().to_crop_blob_rgb(_p(lines), len(lines), _p(pixels), _p(bg3), bg3.shape[1], method, out_w, out_h, _p(out))
    return out


def blob_orientation(lines):
    """pv::Blob::calculate_moments -> (orientation, (cx, cy))."""
    lines = np.ascontiguousarray(lines)
    c = np.zeros(2, np.float32)
    a = lib().to_blob_orientation(_p(lines), len(lines), _p(c))
    return float(np.float32(a)), (float(c[0]), float(c[1]))


def moments_matrix(orientation, bw, bh, out_w=80, out_h=80):
    m = np.zeros(6, np.float64)
    lib().to_moments_matrix(C.c_float(orientation), bw, bh, out_w, out_h, _p(m))
    return m.reshape(2, 3)


def warp_affine_u8(src, M, out_w=80, out_h=80):
    """cv::warpAffine(src, M, (out_w, out_h), INTER_LINEAR, BORDER_CONSTANT) for 8-bit single-channel images."""
    src = np.ascontiguousarray(src, np.uint8); M = np.ascontiguousarray(M, np.float64)
    out = np.zeros((out_h, out_w), np.uint8)
    lib().to_warp_affine_u8(_p(src), src.shape[1], src.shape[0], _p(M), _p(out), out_w, out_h)
    return out


def crop_blob_moments(lines, pixels, bg, method=DIFF_ABSOLUTE, out_w=80, out_h=80):
    """individual_image_normalization = moments (FilterCache.cpp:329-341): rotated by the blob's orientation."""
    lines = np.ascontiguousarray(lines); pixels = np.ascontiguousarray(pixels, np.uint8)
    bg = np.ascontiguousarray(bg, np.uint8)
    out = np.zeros((out_h, out_w), np.uint8)
    lib().to_crop_blob_moments(_p(lines), len(lines), _p(pixels), _p(bg), bg.shape[1], method, out_w, out_h, _p(out))
    return out


def segment_batch(frames, bg, params: Params, crop_method=DIFF_ABSOLUTE, out_w=80, out_h=80,
                  max_crops=0, threads=0):
    """CPU baseline driver: frames over pthreads.  Returns (n_blobs[n], crops or None)."""
    frames = np.ascontiguousarray(frames, np.uint8); bg = np.ascontiguousarray(bg, np.uint8)
    n, h, w = frames.shape
    nb = np.zeros(n, np.int32)
    crops = np.zeros((n, max_crops, out_h, out_w), np.uint8) if max_crops > 0 else None
    pc = params.c()
    lib().to_segment_batch(_p(frames), n, _p(bg), w, h, C.byref(pc), crop_method, out_w, out_h,
                           max_crops, _p(crops), _p(nb), threads)
    return nb, crops


def num_threads() -> int:
    return int(lib().to_num_threads())


def average(frames, method="mean"):
    """AveragingAccumulator restated: frames (n,H,W) u8 -> background (H,W) u8."""
    frames = np.ascontiguousarray(frames, np.uint8)
    n, h, w = frames.shape
    out = np.empty((h, w), np.uint8)
    lib().to_average(_p(frames), n, w, h, {"mean": 0, "mode": 1, "max": 2, "min": 3}[method], _p(out))
    return out


def rethreshold(blobs: "Blobs", bg, threshold: int, method=DIFF_ABSOLUTE, rgb=False, keep_single=False) -> "Blobs":
    """pixel::threshold_blob on every blob of a frame (tracker-side re-threshold, comparison >=).  rgb: the blobs carry
    B,G,R per pixel and bg is the GREY image of the background (Background's cvtColor, Background.cpp:71-77).
    As the tracker's entry does (PixelTree.cpp:344-356), sub-blobs with a payload of one byte (a single grey pixel) are dropped;
    keep_single=True gives every sub-blob (what threshold_get_biggest_blob, the posture loop, chooses from)."""
    bg = np.ascontiguousarray(bg, np.uint8)
    lines = np.ascontiguousarray(blobs.lines); px = np.ascontiguousarray(blobs.pixels, np.uint8)
    lo = np.ascontiguousarray(blobs.line_off, np.int64); po = np.ascontiguousarray(blobs.px_off, np.int64)
    capL = capB = max(16, len(px) + 1); capP = max(16, len(px))
    ol = np.zeros(capL, LINE_DTYPE); op = np.zeros(capP, np.uint8)
    olo = np.zeros(capB + 1, np.int64); opo = np.zeros(capB + 1, np.int64)
    fn = lib().to_rethreshold_frame_rgb if rgb else lib().to_rethreshold_frame
    lib().to_rethreshold_keep_single(int(bool(keep_single)))
    try:
        k = fn(_p(lines), _p(lo), _p(px), _p(po), len(blobs), _p(bg), bg.shape[1], method, int(threshold),
               _p(ol), capL, _p(op), capP, _p(olo), _p(opo), capB)
    finally:
        lib().to_rethreshold_keep_single(0)
    assert k >= 0, k
    return Blobs(ol[:olo[k]].copy(), op[:opo[k]].copy(), olo[:k + 1].copy(), opo[:k + 1].copy())


def find_outer_points(lines):
    """pixel::find_outer_points (C/processing/PixelTree.cpp:497-651): every outline of a blob (list of (n,2) float32 arrays,
    in the order Tree::generate_edges returns them), coordinates relative to the blob's bounding box origin."""
    lines = np.ascontiguousarray(lines)
    npx = int((lines["x1"].astype(np.int64) - lines["x0"] + 1).sum())
    pts = np.zeros((4 * npx + 8, 2), np.float32); off = np.zeros(2 * npx + 8, np.int64)
    k = lib().to_find_outer_points(_p(lines), len(lines), _p(pts), len(pts), _p(off), len(off) - 1)
    if k < 0:
        raise MemoryError
    return [pts[off[i]:off[i + 1]].copy() for i in range(k)]


def longest_outline(lines):
    """The outline calculate_posture selects: the first one of maximal size (T/tracking/Posture.cpp:341-348)."""
    best = None
    for ol in find_outer_points(lines):
        if best is None or len(ol) > len(best):
            best = ol
    return best if best is not None else np.zeros((0, 2), np.float32)


def outline_resample(points, distance=1.0):
    """Outline::resample (T/tracking/Outline.cpp:724-766)."""
    points = np.ascontiguousarray(points, np.float32)
    per = float(np.linalg.norm(np.roll(points, -1, 0) - points, axis=1).sum()) if len(points) else 0.0
    cap = len(points) + 16 + (int(per / distance * 1.01) if distance > 0 else 0)
    out = np.zeros((cap, 2), np.float32)
    n = lib().to_outline_resample(_p(points), len(points), C.c_float(distance), _p(out), cap)
    if n < 0:
        raise MemoryError
    return out[:n].copy()


def blob_recount(lines, pixels, bg, threshold, method, cm_per_pixel=1.0, channels=1):
    """pv::Blob::recount(threshold, background) (C/processing/PVBlob.cpp:934-1027) = raw_recount * SQR(cm_per_pixel), raw_recount the sum over
    the blob's lines of Background::count_above_threshold (C/processing/Background.h:430-489): pixels with  none: v >= T | absolute:
    |bg - v| >= T | sign: bg - v >= T  (int32), v = the pixel's grey value (rgb8 blobs: cmn::bgr2gray, Background.h:76-81) and bg the
    background's grey image; threshold 0 returns num_pixels (:942-949).  The count is accumulated in a float, the product is float."""
    lines = np.ascontiguousarray(lines); px = np.ascontiguousarray(pixels, np.uint8)
    lens = lines["x1"].astype(np.int64) - lines["x0"] + 1
    if threshold == 0:
        return np.float32(np.float32(lens.sum()) * np.float32(np.float32(cm_per_pixel) * np.float32(cm_per_pixel)))
    if channels == 3:
        v = bgr2gray_tracker(px.reshape(-1, 3)).astype(np.int32).ravel()
    else:
        v = px.astype(np.int32)
    rec, o = np.float32(0), 0
    for l, n in zip(lines, lens):
        n = int(n)
        b = bg[int(l["y"]), int(l["x0"]):int(l["x0"]) + n].astype(np.int32)
        vv = v[o:o + n]; o += n
        d = vv if method == DIFF_NONE else (np.abs(b - vv) if method == DIFF_ABSOLUTE else b - vv)
        rec = np.float32(rec + np.float32(int((d >= int(threshold)).sum())))
    return np.float32(rec * np.float32(np.float32(cm_per_pixel) * np.float32(cm_per_pixel)))
