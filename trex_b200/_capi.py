"""ctypes binding of libtrexb200.so (include/trexb200.h).  The library is required: importing the
product without the built CUDA extension fails loudly -- there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtrexb200.so")

TB_OK, TB_ERR_INVALID, TB_ERR_CUDA, TB_ERR_STATE, TB_ERR_CAPACITY = 0, -1, -2, -3, -4


class TrexB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[tb_status {code}] {msg}")
        self.code = code


class SegParams(C.Structure):
    _fields_ = [("detect_threshold", C.c_int32), ("threshold_maximum", C.c_int32),
                ("enable_difference", C.c_int32), ("detect_threshold_is_absolute", C.c_int32),
                ("image_invert", C.c_int32), ("use_closing", C.c_int32), ("closing_size", C.c_int32),
                ("dilation_size", C.c_int32), ("cm_per_pixel", C.c_float), ("n_size_ranges", C.c_int32),
                ("size_lo", C.c_double * 4), ("size_hi", C.c_double * 4),
                ("color_channel", C.c_int32), ("blur_difference", C.c_int32),
                ("use_adaptive_threshold", C.c_int32), ("adaptive_threshold_scale", C.c_float), ("open_size", C.c_int32)]


class SegConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("max_batch", C.c_int32),
                ("max_runs_per_frame", C.c_int32), ("max_pixels_per_frame", C.c_int32),
                ("max_crops_per_frame", C.c_int32), ("crop_width", C.c_int32), ("crop_height", C.c_int32),
                ("crop_method", C.c_int32), ("channels", C.c_int32), ("encoding", C.c_int32),
                ("crop_normalize", C.c_int32), ("crop_scale", C.c_float)]


class FrameInfo(C.Structure):
    _fields_ = [("blob_begin", C.c_uint32), ("n_blobs", C.c_uint32), ("line_begin", C.c_uint32),
                ("n_lines", C.c_uint32), ("px_begin", C.c_uint32), ("n_pixels", C.c_uint32),
                ("n_runs", C.c_uint32), ("status", C.c_uint32)]


class BlobRec(C.Structure):
    _fields_ = [("line_off", C.c_uint32), ("px_off", C.c_uint32), ("n_lines", C.c_uint32), ("n_pixels", C.c_uint32),
                ("x0", C.c_uint16), ("y0", C.c_uint16), ("x1", C.c_uint16), ("y1", C.c_uint16),
                ("bid", C.c_uint32), ("frame", C.c_uint32)]


class BlobView(C.Structure):
    _fields_ = [("info", FrameInfo), ("recs", C.POINTER(BlobRec)), ("lines", C.c_void_p), ("pixels", C.c_void_p)]


class PostureParams(C.Structure):
    _fields_ = [("outline_smooth_samples", C.c_int32), ("outline_smooth_step", C.c_int32), ("outline_approximate", C.c_int32),
                ("outline_curvature_range_ratio", C.c_float), ("midline_walk_offset", C.c_float), ("peak_mode", C.c_int32),
                ("midline_start_with_head", C.c_int32), ("midline_invert", C.c_int32),
                ("midline_resolution", C.c_int32), ("midline_stiff_percentage", C.c_float)]


class PostureRequest(C.Structure):
    _fields_ = [("params", PostureParams), ("outline_resample", C.c_float), ("normalize", C.c_int32), ("fetch", C.c_int32),
                ("median_midline_length_px", C.c_float), ("individual_image_scale", C.c_float),
                ("move_direction_dev", C.c_void_p), ("fix_length_dev", C.c_void_p), ("median_midline_length_dev", C.c_void_p)]


class PostureView(C.Structure):
    _fields_ = [("n_blobs", C.c_uint32), ("midline_resolution", C.c_uint32), ("midlines", C.c_void_p), ("normalized", C.c_void_p),
                ("norm_points", C.c_void_p), ("crop_valid", C.c_void_p), ("outlines", C.c_void_p), ("raw_points", C.c_void_p),
                ("points", C.c_void_p), ("segments", C.c_void_p)]


class MetaLayout(C.Structure):
    _fields_ = [("base", C.c_void_p), ("gather_bytes", C.c_uint64), ("off_infos", C.c_uint64), ("off_top_id", C.c_uint64),
                ("off_top_p", C.c_uint64), ("off_recs", C.c_uint64), ("batch", C.c_uint32), ("kmax", C.c_uint32)]


class ViConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("channels", C.c_int32),
                ("num_classes", C.c_int32), ("max_images", C.c_int32), ("precision", C.c_int32), ("arch", C.c_int32)]


# every symbol include/trexb200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "tb_last_error", "tb_abi_version", "tb_device_count", "tb_seg_default_params", "tb_seg_create",
    "tb_seg_destroy", "tb_seg_set_params", "tb_seg_set_background", "tb_seg_set_background_c", "tb_seg_submit", "tb_seg_submit_device",
    "tb_seg_wait", "tb_seg_result", "tb_seg_totals", "tb_seg_device_results", "tb_seg_crops",
    "tb_seg_debug_binary", "tb_seg_launch_count", "tb_vi_create", "tb_vi_destroy", "tb_vi_set_tensor",
    "tb_vi_commit", "tb_vi_predict", "tb_vi_predict_device", "tb_vi_wait", "tb_vi_launch_count",
    "tb_seg_profile", "tb_seg_kernel_ms", "tb_vi_profile", "tb_vi_kernel_ms", "tb_seg_set_stream",
    "tb_seg_rethreshold", "tb_seg_outlines", "tb_seg_outline_result", "tb_posture_default_params", "tb_seg_midlines", "tb_seg_midline_result",
    "tb_posture_default_request", "tb_seg_posture", "tb_seg_posture_wait", "tb_seg_posture_result", "tb_seg_posture_device", "tb_seg_posture_ms", "tb_seg_posture_thresholded", "tb_seg_recount",
    "tb_vi_set_top1", "tb_avg_create", "tb_avg_destroy", "tb_avg_add", "tb_avg_add_device", "tb_avg_finalize",
    "tb_seg_metadata", "tb_host_alloc", "tb_host_free", "tb_host_register", "tb_host_unregister", "tb_backend",
]

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(trex_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, vpp = C.c_void_p, C.POINTER(C.c_void_p)
    L.tb_last_error.restype = C.c_char_p
    L.tb_seg_default_params.argtypes = [C.POINTER(SegParams)]
    L.tb_seg_create.argtypes = [C.POINTER(SegConfig), vpp]
    L.tb_seg_destroy.argtypes = [vp]; L.tb_seg_destroy.restype = None
    L.tb_seg_set_params.argtypes = [vp, C.POINTER(SegParams)]
    L.tb_seg_set_background.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int64]
    L.tb_seg_set_background_c.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int64]
    L.tb_seg_submit.argtypes = [vp, C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_int]
    L.tb_seg_submit_device.argtypes = [vp, vp, C.c_int, vp, C.c_int]
    L.tb_seg_wait.argtypes = [vp]
    L.tb_seg_result.argtypes = [vp, C.c_int, C.POINTER(BlobView)]
    L.tb_seg_totals.argtypes = [vp, C.POINTER(C.c_uint32 * 4)]
    L.tb_seg_device_results.argtypes = [vp, vpp, vpp, vpp, vpp, vpp]
    L.tb_seg_crops.argtypes = [vp, vpp, vpp, C.POINTER(C.c_uint32)]
    L.tb_seg_debug_binary.argtypes = [vp, vp, vp]
    L.tb_seg_launch_count.argtypes = [vp]; L.tb_seg_launch_count.restype = C.c_uint64
    L.tb_vi_create.argtypes = [C.POINTER(ViConfig), vpp]
    L.tb_vi_destroy.argtypes = [vp]; L.tb_vi_destroy.restype = None
    L.tb_vi_set_tensor.argtypes = [vp, C.c_char_p, vp, C.c_int64]
    L.tb_vi_commit.argtypes = [vp]
    L.tb_vi_predict.argtypes = [vp, vp, C.c_int, vp, vp]
    L.tb_vi_predict_device.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp]
    L.tb_vi_wait.argtypes = [vp]
    L.tb_vi_launch_count.argtypes = [vp]; L.tb_vi_launch_count.restype = C.c_uint64
    L.tb_seg_profile.argtypes = [vp, C.c_int]
    L.tb_seg_kernel_ms.argtypes = [vp, C.POINTER(C.c_double * 3), C.POINTER(C.c_uint64)]
    L.tb_vi_profile.argtypes = [vp, C.c_int]
    L.tb_vi_kernel_ms.argtypes = [vp, C.POINTER(C.c_double * 5), C.POINTER(C.c_uint64)]
    L.tb_seg_set_stream.argtypes = [vp, vp]
    L.tb_seg_rethreshold.argtypes = [vp, vp, C.c_int]
    L.tb_seg_outlines.argtypes = [vp, C.c_float]
    L.tb_seg_outline_result.argtypes = [vp, vpp, vpp, vpp, C.POINTER(C.c_uint32)]
    L.tb_posture_default_params.argtypes = [C.POINTER(PostureParams)]; L.tb_posture_default_params.restype = None
    L.tb_seg_midlines.argtypes = [vp, C.POINTER(PostureParams)]
    L.tb_seg_midline_result.argtypes = [vp, vpp, vpp, vpp, C.POINTER(C.c_uint32)]
    L.tb_posture_default_request.argtypes = [C.POINTER(PostureRequest)]; L.tb_posture_default_request.restype = None
    L.tb_seg_posture.argtypes = [vp, C.POINTER(PostureRequest)]
    L.tb_seg_posture_thresholded.argtypes = [vp, vp, C.POINTER(PostureRequest), C.c_int]
    L.tb_seg_recount.argtypes = [vp, C.c_int, vp, C.c_uint32]
    L.tb_seg_posture_wait.argtypes = [vp]
    L.tb_seg_posture_result.argtypes = [vp, C.POINTER(PostureView)]
    L.tb_seg_posture_device.argtypes = [vp, vpp, vpp, vpp, vpp, vpp, vpp]
    L.tb_seg_posture_ms.argtypes = [vp, C.POINTER(C.c_double * 3), C.POINTER(C.c_uint64)]
    L.tb_vi_set_top1.argtypes = [vp, vp, vp]
    L.tb_avg_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vpp]
    L.tb_avg_destroy.argtypes = [vp]; L.tb_avg_destroy.restype = None
    L.tb_avg_add.argtypes = [vp, vp, C.c_int]
    L.tb_avg_add_device.argtypes = [vp, vp, C.c_int, vp]
    L.tb_avg_finalize.argtypes = [vp, vp]
    L.tb_seg_metadata.argtypes = [vp, C.POINTER(MetaLayout)]
    L.tb_host_alloc.argtypes = [C.c_size_t, vpp]
    L.tb_host_free.argtypes = [vp]
    L.tb_host_register.argtypes = [vp, C.c_size_t]
    L.tb_host_unregister.argtypes = [vp]
    L.trex_b200_register.argtypes = [vp]
    L.tb_backend.restype = vp
    L.tbdbg_umma_shifted_gemm.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp]
    _lib = L
    return L


def check(code: int):
    if code != TB_OK:
        raise TrexB200Error(code, lib().tb_last_error().decode(errors="replace"))
