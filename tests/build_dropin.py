"""Test infrastructure: builds tests/cpp/_dropin/libshim_dropin.so -- INTEGRATION.md's drop-in shim (first ```cpp block, extracted verbatim) compiled against
the reference's own tracker/python/BackgroundSubtraction.h and linked with trex_b200/libtrexb200.so (tests/cpp/shim_dropin.cpp says what surrounds it).
Needs the reference checkout for the headers, so it is built in the authoring container; the .so is git-ignored and travels to the GPU box with the
snapshot.  Returns the library's path, or None when it can neither be built nor found."""
import os
import re
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_SRC = "/root/reference/Application/src"
REF_COMMON = os.path.join(REF_SRC, "commons", "common")
OUT_DIR = os.path.join(HERE, "cpp", "_dropin")
OUT = os.path.join(OUT_DIR, "libshim_dropin.so")
OUT_B = os.path.join(OUT_DIR, "libshim_dropin_b.so")
REF_FILES = [os.path.join(REF_COMMON, "processing", f) for f in ("RawProcessing.cpp", "Background.cpp", "CPULabeling.cpp", "Brototype.cpp", "Source.cpp", "DLList.cpp", "ListCache.cpp", "BlobIdentity.cpp")] + \
            [os.path.join(REF_SRC, "tracker", "core", "SizeFilters.cpp"), os.path.join(REF_COMMON, "video", "AveragingAccumulator.cpp"), os.path.join(REF_COMMON, "misc", "vec2.cpp")]


def snippet() -> str:
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    return re.findall(r"```cpp\n(.*?)```", text, re.S)[0]


def build(force: bool = False):
    have_ref = os.path.exists(os.path.join(REF_SRC, "tracker", "python", "BackgroundSubtraction.h")) and shutil.which("g++") is not None
    lib = os.path.join(ROOT, "trex_b200", "libtrexb200.so")
    if not have_ref or not os.path.exists(lib):
        return OUT if os.path.exists(OUT) else None
    os.makedirs(OUT_DIR, exist_ok=True)
    inc = os.path.join(OUT_DIR, "integration_snippet_1.inc")
    new = snippet()
    if not os.path.exists(inc) or open(inc).read() != new:
        open(inc, "w").write(new)
    srcs = [os.path.join(HERE, "cpp", "shim_dropin.cpp"), os.path.join(ROOT, "oracle", "ref_detect.cpp")]
    deps = srcs + [inc, os.path.abspath(__file__), os.path.join(ROOT, "include", "trexb200.hpp"), os.path.join(ROOT, "include", "trexb200.h")]
    for d in ("ref_stubs", "ref_stubs_detect"):
        for root, _, files in os.walk(os.path.join(ROOT, "oracle", d)):
            deps += [os.path.join(root, f) for f in files]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    cmd = ["g++", "-std=c++23", "-O2", "-fPIC", "-shared", "-Wl,--no-undefined", "-DREF_DETECT",
           "-I", os.path.join(ROOT, "oracle", "ref_stubs_detect"), "-I", os.path.join(ROOT, "oracle", "ref_stubs"), "-I", REF_COMMON, "-I", os.path.join(REF_SRC, "tracker"),
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "cpp"), *srcs, *REF_FILES,
           "-L", os.path.join(ROOT, "trex_b200"), "-ltrexb200", "-Wl,-rpath,$ORIGIN/../../../trex_b200", "-o", OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the drop-in shim failed:\n" + r.stdout + r.stderr)
    shutil.copy(OUT, OUT_B)          # a second copy = a second set of statics (the shim creates its handle once, with the settings of that moment)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
