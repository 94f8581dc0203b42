"""Test infrastructure (oracle/): times the segmentation stage through the REFERENCE'S OWN code -- tracker/python/BackgroundSubtraction.cpp +
commons/common/processing/RawProcessing.cpp + CPULabeling, compiled unmodified into oracle/_ref/libref_detect.so where the reference checkout is
(oracle/build_ref.py build_detect), with every cv:: call forwarded to the real OpenCV in Python's cv2 (oracle/cv_bridge.py).  One thread; the frames are
handed over as BGR like a video source's (cvtColor included, as in TRex); the callback bridge copies every operand once, which is inside the time.
bench.py's cpu_baseline runs this module in a child process:  python -m oracle.ref_detect_leg sample.npz [budget seconds]  -> one JSON line (or `null`)."""
import ctypes as C
import json
import sys
import time

import numpy as np


def measure(bg, frames, budget_s=3.0):
    from oracle import build_ref
    from oracle.cv_bridge import Bridge
    import cv2
    path = build_ref.build_detect()
    if path is None:
        return None
    cv2.setNumThreads(1)
    lib = C.CDLL(path)
    lib.ref_background_subtraction_apply.restype = C.c_int64
    bridge = Bridge(lib)
    sett = dict(enable_difference=1, detect_threshold_is_absolute=1, detect_threshold=15, threshold_maximum=255, use_closing=0, closing_size=3, use_adaptive_threshold=0,
                adaptive_threshold_scale=2.0, dilation_size=0, image_invert=0, tags_enable=0, tags_equalize_hist=0, tags_threshold=15, cm_per_pixel=1.0, blur_difference=0)
    for k, v in sett.items():
        lib.ref_detect_setting(k.encode(), C.c_double(float(v)))
    lib.ref_detect_meta_encoding(0)
    lib.ref_detect_color_channel(-1)
    filt = np.array([10.0, 100000.0])
    lib.ref_detect_size_filter(filt.ctypes.data_as(C.c_void_p), 1)
    h, w = bg.shape
    cap = h * w // 8
    lines = np.zeros((cap, 4), np.uint16); px = np.zeros(cap * 3, np.uint8); lo = np.zeros(cap + 1, np.int64); po = np.zeros(cap + 1, np.int64); fl = np.zeros(cap, np.uint8)
    enc, called = C.c_int32(), C.c_int32()
    p = lambda a: a.ctypes.data_as(C.c_void_p)      # noqa: E731
    bgr = [np.ascontiguousarray(np.repeat(f[:, :, None], 3, axis=2)) for f in frames]
    bgc = np.ascontiguousarray(bg)
    reps, nb, t0 = 0, 0, time.perf_counter()
    while reps < 1 or time.perf_counter() - t0 < budget_s:
        for f in bgr:
            k = lib.ref_background_subtraction_apply(p(f), h, w, 3, p(bgc), 1, p(lines), C.c_int64(cap), p(px), C.c_int64(len(px)), p(lo), p(po), p(fl), C.c_int64(cap),
                                                     C.byref(enc), C.byref(called))
            if k < 0 or bridge.errors:
                return None
            nb += int(k)
        reps += 1
    return {"seg_only_fps_1_thread": reps * len(bgr) / (time.perf_counter() - t0), "blobs_per_frame": nb / (reps * len(bgr)),
            "what": "the reference's own BackgroundSubtraction::apply (compiled unmodified: oracle/_ref/libref_detect.so) with the real OpenCV behind a callback bridge "
                    "(operand copies inside the time), BGR frames, 1 thread; segmentation stage only"}


if __name__ == "__main__":
    d = np.load(sys.argv[1])
    print(json.dumps(measure(d["bg"], list(d["frames"]), float(sys.argv[2]) if len(sys.argv) > 2 else 3.0)))
