#!/usr/bin/env python
"""Device time of the posture chain alone (CUDA events around the kernels, tb_seg_posture_ms): outlines and midlines (+ normalisation)
per batch.  Usage: bench_posture.py [batch=128] [size=1920x1080] [iters=10] [normalize=1]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import trex_b200
from trex_b200.synthetic import BlobWorld

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
W, H = (int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "1920x1080").split("x"))
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
norm = int(sys.argv[4]) if len(sys.argv) > 4 else 1
world = BlobWorld(h=H, w=W, n_blobs=100, seed=1234)
src = world.frames(min(B, 16))
frames = torch.from_numpy(src[np.arange(B) % len(src)]).cuda()
bs = trex_b200.BackgroundSubtraction(world.bg, settings=trex_b200.DetectSettings(), max_batch=B)
stream = torch.cuda.Stream()
for i in range(3):
    bs.apply_device(frames.data_ptr(), B, stream.cuda_stream, fetch=0)
    bs.posture_async(1.0, normalize=bool(norm), fetch=0)
    stream.synchronize()
bs.profile(True)
for i in range(iters):
    bs.apply_device(frames.data_ptr(), B, stream.cuda_stream, fetch=0)
    bs.posture_async(1.0, normalize=bool(norm), fetch=0)
stream.synchronize()
ms, n = bs.posture_ms()
seg, _ = bs.kernel_ms()
bs.posture_async(1.0, normalize=bool(norm), fetch=2)
bs.posture_wait()
r = bs.posture_result()
pts = int((r["outlines"][:, 3]).sum())
print(json.dumps({"size": f"{W}x{H}", "B": B, "blobs": r["n_blobs"], "outline_points": pts, "points_per_blob": pts / max(r["n_blobs"], 1),
                  "outlines_ms": ms["outlines"] / n, "midlines_ms": ms["midlines"] / n, "normalize": norm,
                  "seg_ms": {k: v / n for k, v in seg.items()}}))
