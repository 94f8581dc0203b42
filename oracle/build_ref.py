"""Test infrastructure: builds oracle/_ref/libref_circulargraph.so from the REFERENCE's own source file
/root/reference/Application/src/commons/common/misc/CircularGraph.cpp (compiled where it lies, unmodified) + the C wrapper oracle/ref_circular_graph.cpp,
against the stand-in headers in oracle/ref_stubs/ (TRex's precompiled header needs OpenCV / glaze / cnpy, absent here).  The rest of the reference's
path (Outline.cpp, PixelTree, RawProcessing ...) does not compile without those libraries: see DESIGN.md s6.
The .so is git-ignored, not gpurun-ignored.  Only tests/ load it."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_COMMON = "/root/reference/Application/src/commons/common"
OUT = os.path.join(HERE, "_ref", "libref_circulargraph.so")


def available() -> bool:
    return os.path.exists(os.path.join(REF_COMMON, "misc", "CircularGraph.cpp")) and shutil.which("g++") is not None


def build(force: bool = False):
    """Returns the path of the library, or None when neither the reference checkout nor a prebuilt library is present."""
    if not available():
        return OUT if os.path.exists(OUT) else None
    srcs = [os.path.join(REF_COMMON, "misc", "CircularGraph.cpp"), os.path.join(HERE, "ref_circular_graph.cpp")]
    deps = srcs + [os.path.join(HERE, "ref_stubs", f) for f in ("commons.pc.h", "misc/ranges.h", "misc/Median.h", "misc/Timer.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    # the arithmetic flags of the oracle: no contraction, no fast-math (TRex's own build does not enable fast-math either: CMakeLists.txt)
    cmd = ["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I", os.path.join(HERE, "ref_stubs"), "-I", REF_COMMON, *srcs, "-o", OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the reference's CircularGraph.cpp failed:\n" + r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
