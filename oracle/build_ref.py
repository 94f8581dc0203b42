"""Test infrastructure: builds oracle/_ref/libref_posture.so from the REFERENCE's own source files, compiled where they lie and unmodified:
    tracker/tracking/Outline.cpp                      Outline::resample / smooth / offset_to_middle / calculate_midline, Midline::post_process / normalize / fix_length
    tracker/tracking/Posture.cpp                      posture::calculate_posture(Frame_t, pv::BlobWeakPtr): the threshold loop
    tracker/tracking/FilterCache.cpp                  constraints::diff_image (none / moments / posture / legacy crops), local_midline_length
    commons/common/misc/CircularGraph.cpp             periodic::curvature / eft / ieft / differentiate / find_peaks (+ its fast::cos polynomial)
    commons/common/misc/curve_discussion.cpp, commons/common/gui/Transform.cpp   (linked by Outline.cpp)
    commons/common/processing/PixelTree.cpp           pixel::find_outer_points with pixel::Tree (its threshold_blob half only has to compile)
    commons/common/processing/{CPULabeling,Brototype,Source,DLList,ListCache}.cpp + misc/IllegalVector.h
                                                      CPULabeling::run: extract_lines -> merge_lines (Brototype) -> run_fast
    commons/common/processing/BlobIdentity.cpp + misc/bid.h   pv::blob_bid / pv::bid::from_data: the blob id
    commons/common/processing/Background.{h,cpp} + processing/encoding.h, misc/EnumClass.h, misc/matharray.h, misc/FormatColor.h
                                                      per-pixel difference / is_different / count_above_threshold, cmn::bgr2gray, imageFromLines;
                                                      with it the Background overload of pixel::threshold_blob is the reference's own code as well
plus the C wrappers oracle/ref_outline.cpp, ref_circular_graph.cpp, ref_pixeltree.cpp, ref_labeling.cpp, ref_background.cpp, ref_posture.cpp and ref_filtercache.cpp, against the stand-in headers in oracle/ref_stubs/ (TRex's precompiled
header needs OpenCV / glaze / cnpy, absent here; its settings cache, drawing and tracker headers are irrelevant to the functions under test).
Outline.cpp and Posture.cpp include "DebugDrawing.h" and "Tracker.h" with quotes, which a compiler resolves next to the including file first; they are
therefore compiled through symbolic links in oracle/_ref/overlay/tracking/ (the files themselves stay in the reference checkout), next to placeholders
for those two headers (Outline.h and Posture.h are linked from the reference).  The rest of the reference's path (RawProcessing.cpp, BackgroundSubtraction.cpp, PVBlob.cpp ...) is tied to
OpenCV calls and TRex's image / settings classes and stays restated-only: DESIGN.md s6.
The .so is git-ignored, not gpurun-ignored.  Only tests/ load it."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/Application/src"
REF_COMMON = os.path.join(REF_SRC, "commons", "common")
OUT = os.path.join(HERE, "_ref", "libref_posture.so")
OVERLAY = os.path.join(HERE, "_ref", "overlay", "tracking")
REF_FILES = [os.path.join(REF_SRC, "tracker", "tracking", "Outline.cpp"), os.path.join(REF_SRC, "tracker", "tracking", "Posture.cpp"),
             os.path.join(REF_SRC, "tracker", "tracking", "FilterCache.cpp"),
             os.path.join(REF_COMMON, "misc", "CircularGraph.cpp"), os.path.join(REF_COMMON, "misc", "vec2.cpp"),
             os.path.join(REF_COMMON, "misc", "curve_discussion.cpp"), os.path.join(REF_COMMON, "gui", "Transform.cpp"),
             os.path.join(REF_COMMON, "processing", "PixelTree.cpp")] + \
            [os.path.join(REF_COMMON, "processing", f) for f in ("CPULabeling.cpp", "Brototype.cpp", "Source.cpp", "DLList.cpp", "ListCache.cpp", "Background.cpp", "BlobIdentity.cpp")]


def available() -> bool:
    return all(os.path.exists(f) for f in REF_FILES) and shutil.which("g++") is not None


def _link(src, dst):
    if os.path.islink(dst) or os.path.exists(dst):
        os.remove(dst)
    os.symlink(src, dst)


def build(force: bool = False):
    """Returns the path of the library, or None when neither the reference checkout nor a prebuilt library is present."""
    if not available():
        return OUT if os.path.exists(OUT) else None
    wrappers = [os.path.join(HERE, f) for f in ("ref_outline.cpp", "ref_circular_graph.cpp", "ref_pixeltree.cpp", "ref_labeling.cpp", "ref_background.cpp", "ref_posture.cpp", "ref_filtercache.cpp")]
    stubs = []
    for root, _, files in os.walk(os.path.join(HERE, "ref_stubs")):
        stubs += [os.path.join(root, f) for f in files]
    deps = REF_FILES + wrappers + stubs + [os.path.join(REF_SRC, "tracker", "tracking", "Outline.h"), os.path.join(REF_COMMON, "processing", "PixelTree.h"),
                                            os.path.abspath(__file__)]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(OVERLAY, exist_ok=True)
    _link(REF_FILES[0], os.path.join(OVERLAY, "Outline.cpp"))
    _link(REF_FILES[1], os.path.join(OVERLAY, "Posture.cpp"))
    _link(REF_FILES[2], os.path.join(OVERLAY, "FilterCache.cpp"))
    for h in ("Outline.h", "Posture.h", "FilterCache.h"):                      # the reference's own headers
        _link(os.path.join(REF_SRC, "tracker", "tracking", h), os.path.join(OVERLAY, h))
    for h in ("DebugDrawing.h", "Tracker.h"):                 # placeholders for the two quote-included headers that pull in the whole tracker
        _link(os.path.join(HERE, "ref_stubs", "tracking", h), os.path.join(OVERLAY, h))
    # the arithmetic flags of the oracle: no contraction, no fast-math (TRex's own CMake files do not enable fast-math either)
    cmd = ["g++", "-std=c++23", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wl,--no-undefined",
           "-I", os.path.join(HERE, "ref_stubs"), "-I", REF_COMMON, "-I", os.path.join(REF_SRC, "tracker"),
           os.path.join(OVERLAY, "Outline.cpp"), os.path.join(OVERLAY, "Posture.cpp"), os.path.join(OVERLAY, "FilterCache.cpp"), *REF_FILES[3:], *wrappers, "-o", OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the reference's posture sources failed:\n" + r.stdout + r.stderr)
    return OUT


# ---- the detection build: BackgroundSubtraction::apply / RawProcessing::generate_binary with every cv:: call forwarded to the real OpenCV (Python's cv2) ----
OUT_DETECT = os.path.join(HERE, "_ref", "libref_detect.so")
REF_FILES_DETECT = [os.path.join(REF_SRC, "tracker", "python", "BackgroundSubtraction.cpp"), os.path.join(REF_COMMON, "processing", "RawProcessing.cpp"),
                    os.path.join(REF_SRC, "tracker", "core", "SizeFilters.cpp"), os.path.join(REF_COMMON, "video", "AveragingAccumulator.cpp"), os.path.join(REF_COMMON, "misc", "vec2.cpp")] + \
                   [os.path.join(REF_COMMON, "processing", f) for f in ("CPULabeling.cpp", "Brototype.cpp", "Source.cpp", "DLList.cpp", "ListCache.cpp", "Background.cpp", "BlobIdentity.cpp")]


def build_detect(force: bool = False):
    """oracle/_ref/libref_detect.so: the reference's detection path (tracker/python/BackgroundSubtraction.cpp, commons/common/processing/RawProcessing.cpp,
    the labeling sources, tracker/core/SizeFilters.cpp), compiled unmodified with -DREF_DETECT against oracle/ref_stubs_detect/ + oracle/ref_stubs/ and the
    wrapper oracle/ref_detect.cpp.  Returns the path, or None when neither the checkout nor a prebuilt library is present."""
    if not (all(os.path.exists(f) for f in REF_FILES_DETECT) and shutil.which("g++") is not None):
        return OUT_DETECT if os.path.exists(OUT_DETECT) else None
    wrapper = os.path.join(HERE, "ref_detect.cpp")
    stubs = []
    for d in ("ref_stubs", "ref_stubs_detect"):
        for root, _, files in os.walk(os.path.join(HERE, d)):
            stubs += [os.path.join(root, f) for f in files]
    deps = REF_FILES_DETECT + [wrapper, os.path.abspath(__file__)] + stubs
    if not force and os.path.exists(OUT_DETECT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT_DETECT) for d in deps):
        return OUT_DETECT
    os.makedirs(os.path.dirname(OUT_DETECT), exist_ok=True)
    cmd = ["g++", "-std=c++23", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wl,--no-undefined", "-DREF_DETECT",
           "-I", os.path.join(HERE, "ref_stubs_detect"), "-I", os.path.join(HERE, "ref_stubs"), "-I", REF_COMMON, "-I", os.path.join(REF_SRC, "tracker"),
           *REF_FILES_DETECT, wrapper, "-o", OUT_DETECT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the reference's detection sources failed:\n" + r.stdout + r.stderr)
    return OUT_DETECT


# ---- the reference's own pv::Blob (PVBlob.{h,cpp}) ----
OUT_PVBLOB = os.path.join(HERE, "_ref", "libref_pvblob.so")
REF_FILES_PVBLOB = [os.path.join(REF_COMMON, "processing", f) for f in ("PVBlob.cpp", "BlobIdentity.cpp", "Background.cpp", "PixelTree.cpp", "CPULabeling.cpp", "Brototype.cpp",
                                                                         "Source.cpp", "DLList.cpp", "ListCache.cpp")] + [os.path.join(REF_COMMON, "misc", "vec2.cpp")]


def build_pvblob(force: bool = False):
    """oracle/_ref/libref_pvblob.so: commons/common/processing/PVBlob.cpp with the REAL PVBlob.h (-DREF_REAL_PVBLOB switches the look-alike of oracle/ref_stubs/ off)
    and what it links against, plus the wrapper oracle/ref_pvblob.cpp.  Returns the path, or None when neither the checkout nor a prebuilt library is present."""
    if not (all(os.path.exists(f) for f in REF_FILES_PVBLOB) and shutil.which("g++") is not None):
        return OUT_PVBLOB if os.path.exists(OUT_PVBLOB) else None
    wrapper = os.path.join(HERE, "ref_pvblob.cpp")
    stubs = []
    for d in ("ref_stubs", "ref_stubs_pvblob"):
        for root, _, files in os.walk(os.path.join(HERE, d)):
            stubs += [os.path.join(root, f) for f in files]
    deps = REF_FILES_PVBLOB + [wrapper, os.path.abspath(__file__)] + stubs
    if not force and os.path.exists(OUT_PVBLOB) and all(os.path.getmtime(d) <= os.path.getmtime(OUT_PVBLOB) for d in deps):
        return OUT_PVBLOB
    os.makedirs(os.path.dirname(OUT_PVBLOB), exist_ok=True)
    cmd = ["g++", "-std=c++23", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wl,--no-undefined", "-DREF_REAL_PVBLOB",
           "-I", os.path.join(HERE, "ref_stubs_pvblob"), "-I", os.path.join(HERE, "ref_stubs"), "-I", REF_COMMON, "-I", os.path.join(REF_SRC, "tracker"),
           *REF_FILES_PVBLOB, wrapper, "-o", OUT_PVBLOB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the reference's pv::Blob failed:\n" + r.stdout + r.stderr)
    return OUT_PVBLOB


if __name__ == "__main__":
    print(build(force=True))
    print(build_detect(force=True))
    print(build_pvblob(force=True))
