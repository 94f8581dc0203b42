#!/usr/bin/env python
"""Micro-benchmark of the segmentation kernels alone (K1 seg_rle, K2 ccl_label, K3 blob_emit) on resident
synthetic 1080p batches; prints per-kernel ms and the HBM roofline fraction of K1.
Knobs (env): TB_SEG_FPC (frames per CTA), TB_SEG_NO_TMA=1 (register-streaming K1).
Usage: bench_seg.py [B] [reps] [channels] [gray|rgb8] [none|moments] [WxH] [outlines]   (WxH: frame size, default 1920x1080; 3840x2160 is
BASELINE config 5's segmentation part; channels 3/4: BGR/BGRA frames, cvtColor fused
into K1; moments: the two extra kernels of the normalised crops are timed by the wall clock of the whole batch)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trex_b200  # noqa: E402
from trex_b200.synthetic import BlobWorld  # noqa: E402

from trex_b200.synthetic import to_color  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
CN = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ENC = sys.argv[4] if len(sys.argv) > 4 else "gray"
NORM = sys.argv[5] if len(sys.argv) > 5 else "none"
W, H = (int(v) for v in (sys.argv[6] if len(sys.argv) > 6 else "1920x1080").split("x"))
world = BlobWorld(h=H, w=W, n_blobs=100, seed=1234)
src = world.frames(16)
bg = world.bg
if CN > 1:
    src = to_color(src, seed=1, channels=CN)
    bg3 = to_color(world.bg, seed=2, channels=3)
    bg = bg3 if ENC == "rgb8" else ((3735 * bg3[..., 0].astype(np.int64) + 19235 * bg3[..., 1].astype(np.int64) + 9798 * bg3[..., 2].astype(np.int64) + 16384) >> 15).astype(np.uint8)
dev = torch.device("cuda", 0)
pool = [torch.from_numpy(src[np.random.default_rng(i).permutation(np.arange(B) % 16)]).to(dev) for i in range(4)]
bs = trex_b200.BackgroundSubtraction(bg, settings=trex_b200.DetectSettings(meta_encoding=ENC, individual_image_normalization=NORM), max_batch=B, max_individuals=128, channels=CN)
stream = torch.cuda.Stream(dev)
for i in range(3):
    bs.apply_device(pool[i % 4].data_ptr(), B, stream.cuda_stream)
bs.wait()
bs.profile(True)
for i in range(reps):
    bs.apply_device(pool[i % 4].data_ptr(), B, stream.cuda_stream)
bs.wait()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for i in range(reps):
    bs.apply_device(pool[i % 4].data_ptr(), B, stream.cuda_stream)
bs.wait(); torch.cuda.synchronize()
batch_ms = (time.perf_counter() - t0) / reps * 1e3
ms, n = bs.kernel_ms()
tot = bs.totals()
outline_ms = midline_ms = None
if len(sys.argv) > 7 and sys.argv[7] == "outlines":      # N4 first stage: find_outer_points + resample of every blob (incl. D2H of the points)
    bs.apply_device(pool[0].data_ptr(), B, stream.cuda_stream, fetch=True); bs.wait()
    lib, h = trex_b200._capi.lib(), bs._h
    import ctypes as C
    lib.tb_seg_outlines(h, C.c_float(1.0))
    t0 = time.perf_counter()
    for i in range(5):
        lib.tb_seg_outlines(h, C.c_float(1.0))
    outline_ms = (time.perf_counter() - t0) / 5 * 1e3
    from trex_b200._capi import PostureParams
    P = PostureParams(); lib.tb_posture_default_params(C.byref(P))
    lib.tb_seg_midlines(h, C.byref(P))
    t0 = time.perf_counter()
    for i in range(5):
        lib.tb_seg_midlines(h, C.byref(P))
    midline_ms = (time.perf_counter() - t0) / 5 * 1e3
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
k1 = ms["seg_rle"] / n
alg = B * W * H * CN + 8 * tot[1]
print(json.dumps({"B": B, "size": f"{W}x{H}", "channels": CN, "encoding": ENC, "normalization": NORM, "batch_ms_wall": batch_ms, "fpc": os.environ.get("TB_SEG_FPC"), "no_tma": os.environ.get("TB_SEG_NO_TMA"),
                  "seg_rle_ms": k1, "GBps": alg / k1 / 1e6, "frac": alg / k1 / 1e6 / peak,
                  "ccl_ms": ms["ccl_label"] / n, "emit_ms": ms["blob_emit"] / n, "blobs": tot[0], "outlines_ms_incl_d2h": outline_ms, "midlines_ms_incl_d2h": midline_ms}))
