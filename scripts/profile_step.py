#!/usr/bin/env python
"""A minimal run of the device chain for ncu: `steps` passes of seg -> CNN (-> posture) over one resident batch.
Usage: profile_step.py [batch=64] [precision=fp16c] [steps=2] [posture=0] [size=1920x1080]
  ncu --set full --clock-control none --import-source on -s <launches of the first pass> -c <launches of one pass> -o gpurun_out/step python scripts/profile_step.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import trex_b200
from trex_b200.synthetic import BlobWorld
from trex_b200.weights import random_v118_3_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
precision = sys.argv[2] if len(sys.argv) > 2 else "fp16c"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
posture = int(sys.argv[4]) if len(sys.argv) > 4 else 0
W, H = (int(v) for v in (sys.argv[5] if len(sys.argv) > 5 else "1920x1080").split("x"))
world = BlobWorld(h=H, w=W, n_blobs=100, seed=1234)
src = world.frames(min(B, 16))
frames = torch.from_numpy(src[np.arange(B) % len(src)]).cuda()
bs = trex_b200.BackgroundSubtraction(world.bg, settings=trex_b200.DetectSettings(), max_batch=B, max_individuals=128)
net = trex_b200.VINetwork(100, max_images=B * 128, precision=precision)
net.load_weights(random_v118_3_state_dict(100, seed=0))
crops_p, ncrops_p, _, _, _ = bs.device_results()
probs = torch.empty((B * 128, 100), dtype=torch.float32, device="cuda")
stream = torch.cuda.Stream()
l0 = bs.launch_count() + net.launch_count()
for i in range(steps):
    bs.apply_device(frames.data_ptr(), B, stream.cuda_stream, fetch=0)
    net.predict_device(crops_p, B * 128, ncrops_p, probs.data_ptr(), 0, stream.cuda_stream)
    if posture:
        bs.posture_async(1.0, normalize=True, fetch=0)
    stream.synchronize()
    if i == 0:
        print("launches per pass:", bs.launch_count() + net.launch_count() - l0, flush=True)
bs.wait()
print("blobs", bs.totals())
