"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares; creating a handle without a GPU fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from trex_b200 import _capi
    return _capi


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "trexb200.h")).read()
    declared = set(re.findall(r"\b(tb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(built.SYMBOLS), declared ^ set(built.SYMBOLS)
    L = built.lib()
    for s in declared:
        assert hasattr(L, s), s
    assert hasattr(L, "trex_b200_register")          # the one symbol a host resolves after dlopen (SURVEY s8b A'')
    assert L.tb_abi_version() == 6


def test_struct_sizes_match_header(built):
    assert C.sizeof(built.BlobRec) == 32 and C.sizeof(built.FrameInfo) == 32
    assert C.sizeof(built.SegParams) == 8 * 4 + 4 + 4 + 64 + 16 + 8
    assert C.sizeof(built.SegConfig) == 56 and C.sizeof(built.ViConfig) == 32


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cfg = built.SegConfig(device=0, width=64, height=48, max_batch=1)
    h = C.c_void_p()
    rc = built.lib().tb_seg_create(C.byref(cfg), C.byref(h))
    assert rc == built.TB_ERR_CUDA
    assert b"no CUDA device" in built.lib().tb_last_error()
    vcfg = built.ViConfig(device=0, width=80, height=80, channels=1, num_classes=10, max_images=4, precision=0)
    assert built.lib().tb_vi_create(C.byref(vcfg), C.byref(h)) == built.TB_ERR_CUDA


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "trex_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no CPU oracle", ""), f"{f} references oracle/"


def test_transform_results_mirror_matches_oracle():
    """Host-side mirror of VINetwork::transform_results (VisualIdentification.cpp:808-828) against the oracle's restatement."""
    import numpy as np
    from oracle import vi
    from trex_b200.visual_identification import VINetwork
    rng = np.random.default_rng(0)
    rows = {i: rng.random(5).astype(np.float32) for i in range(7)}
    for idx in ([0, 1, 2, 3, 4, 5, 6], [0, 2, 3], [1, 4], [3], []):
        a = VINetwork.transform_results(7, idx, rows, 5)
        b = vi.transform_results(7, idx, rows, 5)
        assert np.array_equal(a, b), idx
    assert (VINetwork.transform_results(4, [2], rows, 5)[:2] == -1).all()


def test_blob_properties_mirror():
    """Blob.num_pixels / Blob.center follow pv::Blob::calculate_properties (PVBlob.cpp:216-243)."""
    import numpy as np
    from trex_b200.background_subtraction import LINE_DTYPE, Blob
    lines = np.array([(10, 14, 7, 0), (8, 20, 8, 0), (9, 9, 9, 0)], LINE_DTYPE)
    b = Blob(lines, np.zeros(19, np.uint8), 0, (8, 7, 20, 9))
    assert b.num_pixels == 19 and b.center == (8 + 13 * 0.5, 7 + 3 * 0.5)
