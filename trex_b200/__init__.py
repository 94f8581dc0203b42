"""trex_b200: TRex's segmentation + identification hot path on NVIDIA B200 (sm_100a).

Host-side mirror of the reference interfaces for this path, over the C ABI of include/trexb200.h:
    BackgroundSubtraction   Application/src/tracker/python/BackgroundSubtraction.{h,cpp}
    VINetwork               Application/src/tracker/ml/VisualIdentification.{h,cpp}
The CUDA library is mandatory; nothing on the path computes on the CPU (the .pv container reader / writer are file plumbing).
"""
from ._capi import LIB_PATH, TrexB200Error, lib  # noqa: F401
from .background_subtraction import BackgroundSubtraction, Blob, DetectSettings  # noqa: F401
from .visual_identification import VINetwork  # noqa: F401
from .averaging import AveragingAccumulator  # noqa: F401
from .pv_writer import PVWriter  # noqa: F401
from .pv_reader import PVReader  # noqa: F401

__all__ = ["AveragingAccumulator", "PVWriter", "PVReader", "BackgroundSubtraction", "Blob", "DetectSettings", "VINetwork", "TrexB200Error", "lib", "LIB_PATH"]
