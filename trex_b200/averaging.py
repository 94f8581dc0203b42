"""Host mirror of `cmn::AveragingAccumulator` (Application/src/commons/common/video/AveragingAccumulator.{h,cpp}):
add frames, finalize() -> background image for BackgroundSubtraction.set_background.  Runs on the GPU through
the tb_avg_* C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._capi import check, lib

METHODS = {"mean": 0, "mode": 1, "max": 2, "min": 3}      # averaging_method_t (AveragingAccumulator.h:6)


def sample_indices(num_frames: int, samples: int, start: int = 0):
    """Frames VideoSource::generate_average feeds to the accumulator (C/video/VideoSource.cpp:1040-1060):
    `samples` indices spread over [start, start+num_frames), offset_i = round(i * max(1, (num_frames-1)/samples))."""
    if samples <= 1:
        return [start]
    samples = min(samples, num_frames)
    step = max(1.0, (num_frames - 1) / float(samples))
    out = []
    for i in range(samples):
        off = int(i * step + 0.5)                       # std::round for non-negative values
        out.append(start + min(off, num_frames - 1))
    return out


class AveragingAccumulator:
    def __init__(self, width: int, height: int, mode: str = "mean", device: int = 0):
        self.width, self.height, self.mode = int(width), int(height), mode
        self._h = C.c_void_p()
        check(lib().tb_avg_create(device, self.width, self.height, METHODS[mode], C.byref(self._h)))

    def add(self, frames):
        """One HxW frame or an (n,H,W) stack of u8 frames (AveragingAccumulator::add)."""
        f = np.ascontiguousarray(frames, np.uint8)
        if f.ndim == 2:
            f = f[None]
        assert f.shape[1:] == (self.height, self.width), f.shape
        check(lib().tb_avg_add(self._h, f.ctypes.data_as(C.c_void_p), f.shape[0]))

    def add_device(self, frames_ptr: int, n: int, stream: int = 0):
        check(lib().tb_avg_add_device(self._h, C.c_void_p(frames_ptr), n, C.c_void_p(stream)))

    def finalize(self) -> np.ndarray:
        out = np.empty((self.height, self.width), np.uint8)
        check(lib().tb_avg_finalize(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def deinit(self):
        if self._h:
            lib().tb_avg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.deinit()
        except Exception:
            pass
