"""Host mirror of `track::BackgroundSubtraction` (Application/src/tracker/python/BackgroundSubtraction.h:10-27)
on top of the tb_seg_* C ABI.  Same call shape as the reference class:

    BackgroundSubtraction(average)          ctor with optional background      (.cpp:50-84)
    set_background(average)                 un-pauses the pipeline             (.cpp:86-90)
    apply(frames) -> list[list[Blob]]       one blob list per frame            (.cpp:126-347)
    fps()                                   mean of batch / elapsed            (.cpp:21-35,345)
    deinit()

Settings are the reference's keys (DetectSettings), read where the reference reads them
(RawProcessing.cpp:266-327, BackgroundSubtraction.cpp:137-139).  Errors: the reference fails the
tile's promise with the exception; here the exception propagates to the caller of apply().
"""
from __future__ import annotations

import ctypes as C
import time
from dataclasses import dataclass, field

import numpy as np

from . import _capi
from ._capi import BlobView, SegConfig, SegParams, check, lib

LINE_DTYPE = np.dtype([("x0", "<u2"), ("x1", "<u2"), ("y", "<u2"), ("pad", "<u2")])
REC_DTYPE = np.dtype([("line_off", "<u4"), ("px_off", "<u4"), ("n_lines", "<u4"), ("n_pixels", "<u4"),
                      ("x0", "<u2"), ("y0", "<u2"), ("x1", "<u2"), ("y1", "<u2"), ("bid", "<u4"), ("frame", "<u4")])
NORM_DTYPE = np.dtype([("len", "<f4"), ("angle", "<f4"), ("offx", "<f4"), ("offy", "<f4"), ("n_points", "<u4"), ("flags", "<u4"),
                       ("tail", "<i4"), ("head", "<i4")])
INFO_DTYPE = np.dtype([("blob_begin", "<u4"), ("n_blobs", "<u4"), ("line_begin", "<u4"), ("n_lines", "<u4"),
                       ("px_begin", "<u4"), ("n_pixels", "<u4"), ("n_runs", "<u4"), ("status", "<u4")])


@dataclass
class DetectSettings:
    """GlobalSettings keys of the path with the reference defaults."""
    detect_threshold: int = 15
    threshold_maximum: int = 255
    enable_difference: bool = True
    detect_threshold_is_absolute: bool = True
    image_invert: bool = False
    use_closing: bool = False
    closing_size: int = 3
    dilation_size: int = 0
    blur_difference: bool = False              # grabber default_config.cpp:125
    use_adaptive_threshold: bool = False       # T/core/default_config.cpp:1162
    adaptive_threshold_scale: float = 2.0      # :1161
    cm_per_pixel: float = 1.0
    detect_size_filter: list = field(default_factory=lambda: [(10.0, 100000.0)])
    individual_image_size: tuple = (80, 80)
    # crop content: track_background_subtraction + track_threshold_is_absolute (FilterCache.cpp:165-174)
    track_background_subtraction: bool = True
    track_threshold_is_absolute: bool = True
    # colour handling of BackgroundSubtraction::apply (.cpp:151-188): meta_encoding "gray" | "rgb8", color_channel
    meta_encoding: str = "gray"        # "gray" | "rgb8" | "r3g3b2"
    color_channel: int | None = None
    # individual_image_normalization (FilterCache.cpp:318-346): "none" | "moments" | "posture" | "legacy" (the last two: the crops are
    # re-rendered by posture(); until then they are the "none" crops)
    individual_image_normalization: str = "none"
    individual_image_scale: float = 1.0        # T/core/default_config.cpp; != 1: resize_image (INTER_NEAREST) before the pad / crop
    open_size: int = 0                         # NOT a reference key: north_star's optional n x n open of the threshold mask (cv2 MORPH_OPEN, ones(n,n)); 0 = off

    def c_params(self) -> SegParams:
        p = SegParams()
        lib().tb_seg_default_params(C.byref(p))
        p.detect_threshold = int(self.detect_threshold); p.threshold_maximum = int(self.threshold_maximum)
        p.enable_difference = int(self.enable_difference)
        p.detect_threshold_is_absolute = int(self.detect_threshold_is_absolute)
        p.image_invert = int(self.image_invert); p.use_closing = int(self.use_closing)
        p.closing_size = int(self.closing_size); p.dilation_size = int(self.dilation_size)
        p.cm_per_pixel = float(self.cm_per_pixel)
        p.n_size_ranges = len(self.detect_size_filter)
        for i, (lo, hi) in enumerate(self.detect_size_filter[:4]):
            p.size_lo[i], p.size_hi[i] = float(lo), float(hi)
        p.color_channel = -1 if self.color_channel is None else int(self.color_channel)
        p.blur_difference = int(self.blur_difference); p.use_adaptive_threshold = int(self.use_adaptive_threshold)
        p.adaptive_threshold_scale = float(self.adaptive_threshold_scale)
        p.open_size = int(self.open_size)
        return p

    @property
    def crop_method(self) -> int:
        if not self.track_background_subtraction:
            return 0
        return 1 if self.track_threshold_is_absolute else 2


@dataclass
class Blob:
    """blob::Pair (C/misc/types.h:591-604) + the geometric id pv::bid (C/misc/bid.h:87-94)."""
    lines: np.ndarray      # LINE_DTYPE
    pixels: np.ndarray     # uint8
    bid: int
    bounds: tuple          # x0, y0, x1, y1 inclusive

    @property
    def num_pixels(self) -> int:
        """pv::Blob::calculate_properties (C/processing/PVBlob.cpp:216-243): the sum of the line lengths."""
        return int((self.lines["x1"].astype(np.int64) - self.lines["x0"] + 1).sum())

    @property
    def center(self) -> tuple:
        """_properties.center = bounds.pos() + bounds.size() * 0.5 (PVBlob.cpp:239), bounds = (x, y, maxx - x + 1, maxy - y + 1)."""
        x0, y0, x1, y1 = self.bounds
        return (x0 + (x1 - x0 + 1) * 0.5, y0 + (y1 - y0 + 1) * 0.5)


class BackgroundSubtraction:
    def __init__(self, average: np.ndarray | None = None, *, width=None, height=None, settings: DetectSettings | None = None,
                 max_batch=16, max_individuals=0, device=0, max_runs_per_frame=0, max_pixels_per_frame=0, channels=1):
        """channels: bytes per pixel of the frames apply() receives (1 gray, 3 BGR, 4 BGRA; TileImage.images[0])."""
        if average is not None:
            height, width = average.shape[:2]
        if width is None or height is None:
            raise ValueError("BackgroundSubtraction needs an average image or width/height")
        self.settings = settings or DetectSettings()
        if self.settings.meta_encoding not in ("gray", "rgb8", "r3g3b2"):
            raise _capi.TrexB200Error(_capi.TB_ERR_INVALID, f"meta_encoding {self.settings.meta_encoding!r} is not built (gray, rgb8, r3g3b2)")
        self.width, self.height, self.max_batch = int(width), int(height), int(max_batch)
        self.channels = int(channels)
        self.out_channels = 3 if self.settings.meta_encoding == "rgb8" else 1           # bytes per blob pixel
        self.crop_channels = 1 if self.settings.meta_encoding == "gray" else 3           # r3g3b2 blobs render as B,G,R (Background.cpp:134-139)
        cfg = SegConfig(device=device, width=self.width, height=self.height, max_batch=self.max_batch,
                        max_runs_per_frame=max_runs_per_frame, max_pixels_per_frame=max_pixels_per_frame,
                        max_crops_per_frame=int(max_individuals),
                        crop_width=self.settings.individual_image_size[0],
                        crop_height=self.settings.individual_image_size[1],
                        crop_method=self.settings.crop_method, channels=self.channels,
                        encoding={"gray": 0, "rgb8": 1, "r3g3b2": 2}[self.settings.meta_encoding],
                        crop_normalize={"none": 0, "moments": 1, "posture": 2, "legacy": 3}[self.settings.individual_image_normalization],
                        crop_scale=float(self.settings.individual_image_scale))
        self.max_individuals = int(max_individuals)
        self._h = C.c_void_p()
        check(lib().tb_seg_create(C.byref(cfg), C.byref(self._h)))
        self.update_settings(self.settings)
        self._time = 0.0
        self._samples = 0.0
        self._has_background = False
        self._last_n = 0
        if average is not None:
            self.set_background(average)

    # -- reference interface -------------------------------------------------------------------
    def set_background(self, average: np.ndarray):
        a = np.ascontiguousarray(average, np.uint8)
        if a.ndim == 3 and a.shape[2] == 1:
            a = a[..., 0]
        ch = 1 if a.ndim == 2 else a.shape[2]
        check(lib().tb_seg_set_background_c(self._h, a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[0], ch, a.strides[0]))
        self._has_background = True

    def update_settings(self, settings: DetectSettings):
        self.settings = settings
        p = settings.c_params()
        check(lib().tb_seg_set_params(self._h, C.byref(p)))

    def apply(self, frames, fetch=True, fetch_crops=None, materialize=True):
        """frames: sequence of HxW uint8 arrays (or one (n,H,W) array).  Returns one list of Blob per frame
        (materialize=False: results stay in the pinned result buffers, see raw_result / totals)."""
        if fetch_crops is None:
            fetch_crops = self.max_individuals > 0
        level = 0 if not fetch else (2 if fetch_crops else 1)
        t0 = time.perf_counter()
        frames = [np.ascontiguousarray(f, np.uint8) for f in frames]
        out = []
        for i in range(0, len(frames), self.max_batch):
            chunk = frames[i:i + self.max_batch]
            ptrs = (C.c_void_p * len(chunk))(*[f.ctypes.data_as(C.c_void_p) for f in chunk])
            want = (self.height, self.width) if self.channels == 1 else (self.height, self.width, self.channels)
            for f in chunk:
                if f.shape != want:
                    raise _capi.TrexB200Error(_capi.TB_ERR_INVALID, f"frame shape {f.shape} != {want}")
            check(lib().tb_seg_submit(self._h, ptrs, len(chunk), self.width * self.channels, level))
            self._last_n = len(chunk)
            check(lib().tb_seg_wait(self._h))
            if fetch and materialize:
                out.extend(self.result(j) for j in range(len(chunk)))
        dt = time.perf_counter() - t0
        if frames:
            self._time += len(frames) / max(dt, 1e-12)
            self._samples += 1
        return out

    def set_stream(self, stream: int):
        """cudaStream_t handle used by submit()/apply() (0 = the handle's private stream)."""
        check(lib().tb_seg_set_stream(self._h, C.c_void_p(stream)))

    def submit(self, frames, fetch=1):
        """Asynchronous half of apply(): enqueue H2D + kernels for up to max_batch frames ((n,H,W) u8 array,
        ideally pinned).  fetch: 0 headers only, 1 blobs, 2 blobs + crops.  Pair with wait()."""
        n = len(frames)
        base, stride = frames.ctypes.data, frames.strides[0]
        ptrs = (C.c_void_p * n)(*[base + i * stride for i in range(n)])
        check(lib().tb_seg_submit(self._h, ptrs, n, frames.strides[1], int(fetch)))
        self._last_n = n

    def apply_device(self, frames_ptr: int, n: int, stream: int = 0, fetch=False):
        """Frames already resident in HBM (n packed HxW u8); work is ordered on `stream`."""
        check(lib().tb_seg_submit_device(self._h, C.c_void_p(frames_ptr), n, C.c_void_p(stream), int(fetch)))
        self._last_n = n

    def rethreshold(self, tracker: "BackgroundSubtraction", fetch=1, materialize=True):
        """pixel::threshold_blob on every blob of this handle's last batch (PixelTree.cpp:186-291).  `tracker` is a
        second BackgroundSubtraction of the same size whose settings hold the tracker keys (detect_threshold :=
        track_threshold with comparison >=, enable_difference := track_background_subtraction, ...).  Returns one
        list of Blob per frame (tracker-side blobs) when fetch and materialize are set."""
        check(lib().tb_seg_rethreshold(self._h, tracker._h, int(fetch)))
        check(lib().tb_seg_wait(tracker._h))
        if fetch and materialize:
            return [tracker.result(i) for i in range(self._last_n)]
        return None

    def recount(self, threshold: int) -> np.ndarray:
        """pv::Blob::recount(threshold, background) for every blob of the last batch (PVBlob.cpp:934-1027): pixels whose difference
        (this handle's method) is >= threshold, times SQR(cm_per_pixel); float32 per blob in batch order."""
        n = self.totals()[0]
        out = np.zeros(max(n, 1), np.float32)
        check(lib().tb_seg_recount(self._h, int(threshold), out.ctypes.data_as(C.c_void_p), len(out)))
        return out[:n]

    def outlines(self, outline_resample=1.0):
        """pixel::find_outer_points + the outline calculate_posture selects + Outline::resample for every blob of the last
        batch (PixelTree.cpp:497-651, Posture.cpp:341-348, Outline.cpp:724-766).  Returns (raw, resampled): two lists with
        one (n,2) float32 array per blob of the batch (blob order of the batch), coordinates relative to the blob's bounds."""
        check(lib().tb_seg_outlines(self._h, C.c_float(outline_resample)))
        rp, pp, qp, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint32()
        check(lib().tb_seg_outline_result(self._h, C.byref(rp), C.byref(pp), C.byref(qp), C.byref(n)))
        if n.value == 0:
            return [], []
        recs = np.ctypeslib.as_array(C.cast(rp, C.POINTER(C.c_uint32)), (n.value, 4))
        n_raw, n_res = int((recs[:, 0] + recs[:, 1]).max()), int((recs[:, 2] + recs[:, 3]).max())
        raw = np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_float)), (max(n_raw, 1), 2))
        res = np.ctypeslib.as_array(C.cast(qp, C.POINTER(C.c_float)), (max(n_res, 1), 2))
        return ([raw[o:o + k].copy() for o, k in recs[:, :2]], [res[o:o + k].copy() for o, k in recs[:, 2:]])

    def midlines(self, outline_resample=1.0, **posture_settings):
        """Outline::calculate_midline for every blob of the last batch (T/tracking/Outline.cpp:768-868): runs
        outlines(outline_resample) first.  posture_settings: fields of tb_posture_params (outline_smooth_samples, ...).
        Returns one (segments (n,4) float32, tail index, head index, outline points (m,2)) per blob; segments is empty where
        the reference reports too few midline segments."""
        from ._capi import PostureParams
        check(lib().tb_seg_outlines(self._h, C.c_float(outline_resample)))
        P = PostureParams()
        lib().tb_posture_default_params(C.byref(P))
        for k, v in posture_settings.items():
            setattr(P, k, v)
        check(lib().tb_seg_midlines(self._h, C.byref(P)))
        orp, a, b, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint32()
        check(lib().tb_seg_outline_result(self._h, C.byref(orp), C.byref(a), C.byref(b), C.byref(n)))
        mrp, pp, sp, m = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint32()
        check(lib().tb_seg_midline_result(self._h, C.byref(mrp), C.byref(pp), C.byref(sp), C.byref(m)))
        if m.value == 0:
            return []
        orecs = np.ctypeslib.as_array(C.cast(orp, C.POINTER(C.c_uint32)), (n.value, 4))
        mrecs = np.ctypeslib.as_array(C.cast(mrp, C.POINTER(C.c_int32)), (m.value, 4))
        top = int((orecs[:, 2] + orecs[:, 3]).max())
        pts = np.ctypeslib.as_array(C.cast(pp, C.POINTER(C.c_float)), (max(top, 1), 2))
        segs = np.ctypeslib.as_array(C.cast(sp, C.POINTER(C.c_float)), (max(top, 1), 4))
        out = []
        for k in range(m.value):
            so, ns, tail, head = (int(v) for v in mrecs[k])
            ro, rn = int(orecs[k, 2]), int(orecs[k, 3])
            out.append((segs[so:so + ns].copy(), tail, head, pts[ro:ro + rn].copy()))
        return out

    def posture_async(self, outline_resample=1.0, normalize=True, fetch=1, median_midline_length_px=0.0, individual_image_scale=1.0,
                      move_direction_dev=0, fix_length_dev=0, median_midline_length_dev=0, **posture_settings):
        """tb_seg_posture: enqueue outlines -> midlines -> Midline::post_process / normalize (-> posture crops) behind the last batch's
        kernels; may be called before wait().  Pair with posture_wait() / posture_result()."""
        from ._capi import PostureRequest
        q = PostureRequest()
        lib().tb_posture_default_request(C.byref(q))
        for k, v in posture_settings.items():
            setattr(q.params, k, v)
        q.outline_resample = float(outline_resample); q.normalize = int(bool(normalize)); q.fetch = int(fetch)
        q.median_midline_length_px = float(median_midline_length_px); q.individual_image_scale = float(individual_image_scale)
        q.move_direction_dev = move_direction_dev or None; q.fix_length_dev = fix_length_dev or None
        q.median_midline_length_dev = median_midline_length_dev or None
        check(lib().tb_seg_posture(self._h, C.byref(q)))

    def _posture_request(self, outline_resample, normalize, fetch, median_midline_length_px, individual_image_scale, posture_settings):
        from ._capi import PostureRequest
        q = PostureRequest()
        lib().tb_posture_default_request(C.byref(q))
        for k, v in posture_settings.items():
            setattr(q.params, k, v)
        q.outline_resample = float(outline_resample); q.normalize = int(bool(normalize)); q.fetch = int(fetch)
        q.median_midline_length_px = float(median_midline_length_px); q.individual_image_scale = float(individual_image_scale)
        return q

    def posture_thresholded(self, posture_handle: "BackgroundSubtraction", track_posture_threshold=0, outline_resample=1.0, normalize=True, fetch=2,
                            median_midline_length_px=0.0, individual_image_scale=1.0, **posture_settings):
        """posture::calculate_posture (Posture.cpp:305-400) for every blob of this handle's last batch: re-threshold at
        track_posture_threshold (+2 per retry) on `posture_handle` (configured like rethreshold()'s tracker handle), biggest sub-blob,
        outline in this blob's frame, midline.  Returns (rounds, posture_handle.posture_result())."""
        q = self._posture_request(outline_resample, normalize, fetch, median_midline_length_px, individual_image_scale, posture_settings)
        rounds = lib().tb_seg_posture_thresholded(self._h, posture_handle._h, C.byref(q), int(track_posture_threshold))
        if rounds < 0:
            check(rounds)
        check(lib().tb_seg_posture_wait(posture_handle._h))
        res = posture_handle.posture_result(n_crops=self.totals()[3])
        return rounds, res

    def posture_wait(self):
        check(lib().tb_seg_posture_wait(self._h))

    def posture_result(self, n_crops=None):
        """dict of numpy views of the last posture call (valid until the next one): midlines (n,4) int32 [seg_off, n_seg, tail, head],
        normalized (structured: len, angle, offx, offy, n_points, flags, tail, head), norm_points (n, resolution, 4), crop_valid,
        and with fetch=2 outlines (n,4) uint32, raw_points, points, segments arenas."""
        from ._capi import PostureView
        v = PostureView()
        check(lib().tb_seg_posture_result(self._h, C.byref(v)))
        n, res = int(v.n_blobs), int(v.midline_resolution)
        out = {"n_blobs": n, "midline_resolution": res}
        if n == 0:
            return out
        if v.midlines:
            out["midlines"] = np.ctypeslib.as_array(C.cast(v.midlines, C.POINTER(C.c_int32)), (n, 4))
        if v.normalized:
            out["normalized"] = np.ctypeslib.as_array(C.cast(v.normalized, C.POINTER(C.c_uint8)), (n * 32,)).view(NORM_DTYPE)
            out["norm_points"] = np.ctypeslib.as_array(C.cast(v.norm_points, C.POINTER(C.c_float)), (n, res, 4))
        if v.crop_valid:
            out["crop_valid"] = np.ctypeslib.as_array(C.cast(v.crop_valid, C.POINTER(C.c_uint8)), (max(self.totals()[3] if n_crops is None else n_crops, 1),))
        if v.outlines:
            orecs = np.ctypeslib.as_array(C.cast(v.outlines, C.POINTER(C.c_uint32)), (n, 4))
            out["outlines"] = orecs
            n_raw, n_res = int((orecs[:, 0] + orecs[:, 1]).max()), int((orecs[:, 2] + orecs[:, 3]).max())
            out["raw_points"] = np.ctypeslib.as_array(C.cast(v.raw_points, C.POINTER(C.c_float)), (max(n_raw, 1), 2))
            out["points"] = np.ctypeslib.as_array(C.cast(v.points, C.POINTER(C.c_float)), (max(n_res, 1), 2))
            out["segments"] = np.ctypeslib.as_array(C.cast(v.segments, C.POINTER(C.c_float)), (max(n_res, 1), 4))
        return out

    def posture(self, outline_resample=1.0, **kw):
        """Synchronous form: posture_async(fetch=2) + posture_wait() + posture_result()."""
        kw.setdefault("fetch", 2)
        self.posture_async(outline_resample, **kw)
        self.posture_wait()
        return self.posture_result()

    def posture_ms(self):
        ms, n = (C.c_double * 3)(), C.c_uint64()
        check(lib().tb_seg_posture_ms(self._h, C.byref(ms), C.byref(n)))
        return dict(zip(("outlines", "midlines", "posture_crops"), ms)), int(n.value)

    def wait(self):
        check(lib().tb_seg_wait(self._h))

    def fps(self) -> float:
        return self._time / self._samples if self._samples else 0.0

    def deinit(self):
        if self._h:
            lib().tb_seg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.deinit()
        except Exception:      # interpreter shutdown
            pass

    # -- results -------------------------------------------------------------------------------
    def raw_result(self, i):
        """(info, recs, lines, pixels) numpy views of frame i (valid until the next apply)."""
        v = BlobView()
        check(lib().tb_seg_result(self._h, i, C.byref(v)))
        nb, nl, npx = v.info.n_blobs, v.info.n_lines, v.info.n_pixels
        recs = np.ctypeslib.as_array(C.cast(v.recs, C.POINTER(C.c_uint8)), (nb * 32,)).view(REC_DTYPE) if nb else np.zeros(0, REC_DTYPE)
        lines = np.ctypeslib.as_array(C.cast(v.lines, C.POINTER(C.c_uint8)), (nl * 8,)).view(LINE_DTYPE) if nl else np.zeros(0, LINE_DTYPE)
        px = np.ctypeslib.as_array(C.cast(v.pixels, C.POINTER(C.c_uint8)), (npx,)) if npx else np.zeros(0, np.uint8)
        return v.info, recs, lines, px

    def result(self, i):
        info, recs, lines, px = self.raw_result(i)
        blobs = []
        for r in recs:
            lo, po = int(r["line_off"]) - info.line_begin, int(r["px_off"]) - info.px_begin
            blobs.append(Blob(lines[lo:lo + int(r["n_lines"])].copy(), px[po:po + int(r["n_pixels"]) * self.out_channels].copy(), int(r["bid"]),
                              (int(r["x0"]), int(r["y0"]), int(r["x1"]), int(r["y1"]))))
        return blobs

    def frame_info(self, i):
        return self.raw_result(i)[0]

    def totals(self):
        t = (C.c_uint32 * 4)()
        check(lib().tb_seg_totals(self._h, C.byref(t)))
        return tuple(int(x) for x in t)

    def crops(self):
        """(crops[n,h,w] u8, blob_index[n]) of the last fetched batch."""
        cp, ip, n = C.c_void_p(), C.c_void_p(), C.c_uint32()
        check(lib().tb_seg_crops(self._h, C.byref(cp), C.byref(ip), C.byref(n)))
        w, h = self.settings.individual_image_size
        shape = (h, w) if self.crop_channels == 1 else (h, w, 3)
        if n.value == 0:
            return np.zeros((0,) + shape, np.uint8), np.zeros(0, np.uint32)
        crops = np.ctypeslib.as_array(C.cast(cp, C.POINTER(C.c_uint8)), (n.value,) + shape).copy()
        idx = np.ctypeslib.as_array(C.cast(ip, C.POINTER(C.c_uint32)), (n.value,)).copy()
        return crops, idx

    def device_results(self):
        """Device pointers (ints): crops, n_crops, crop_blob_index, recs, infos."""
        ps = [C.c_void_p() for _ in range(5)]
        check(lib().tb_seg_device_results(self._h, *[C.byref(p) for p in ps]))
        return tuple(p.value for p in ps)

    def metadata(self):
        """tb_meta_layout of the handle: the device block that holds headers | top-1 ids | top-1 probabilities | blob records in
        the layout of the multi-GPU all-gather (no packing step).  Pass top1_ptrs() to VINetwork.set_top1."""
        m = _capi.MetaLayout()
        check(lib().tb_seg_metadata(self._h, C.byref(m)))
        return m

    def top1_ptrs(self):
        m = self.metadata()
        return m.base + m.off_top_id, m.base + m.off_top_p

    def debug_binary(self, frame: np.ndarray) -> np.ndarray:
        f = np.ascontiguousarray(frame, np.uint8)
        out = np.empty((self.height, self.width) if self.out_channels == 1 else (self.height, self.width, 3), np.uint8)
        check(lib().tb_seg_debug_binary(self._h, f.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
        return out

    def launch_count(self) -> int:
        return int(lib().tb_seg_launch_count(self._h))

    def profile(self, enable=True):
        check(lib().tb_seg_profile(self._h, int(enable)))

    def kernel_ms(self):
        """({kernel: summed ms}, n_batches) since the last call (synchronises)."""
        ms, n = (C.c_double * 3)(), C.c_uint64()
        check(lib().tb_seg_kernel_ms(self._h, C.byref(ms), C.byref(n)))
        return dict(zip(("seg_rle", "ccl_label", "blob_emit"), ms)), int(n.value)
