#!/usr/bin/env python
"""Throughput of the identification networks (visual_identification_version) on resident crops: crops/s and algorithmic
TFLOP/s per network.  v118_3 runs on tcgen05 tensor cores (fp16 / bf16x3), the others on fp32 CUDA cores (vi_nets.cu).
Usage: bench_nets.py [n_crops] [reps]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trex_b200  # noqa: E402
from trex_b200.weights import random_state_dict  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
M = 100
# multiply-accumulates per crop (80x80x1 input, M classes)
MACS = {
    "v118_3": 6400 * 25 * 16 + 1600 * 25 * 16 * 64 + 400 * 25 * 64 * 128 + 12800 * 100 + 100 * M,
    "v100": 6400 * 25 * 16 + 1600 * 25 * 16 * 64 + 400 * 25 * 64 * 100 + 10000 * 100 + 100 * M,
    "v110": 6400 * 25 * 16 + 1600 * 25 * 16 * 64 + 400 * 25 * 64 * 100 + 10000 * 100 + 100 * M,
    "v119": 6400 * 25 * 256 + 1600 * 25 * 256 * 128 + 400 * 25 * 128 * 32 + 100 * 25 * 32 * 128 + 3200 * 1024 + 1024 * M,
    "v200": 6400 * 9 * 64 + 6400 * 9 * 64 * 128 + 676 * 9 * 128 * 256 + 676 * 9 * 256 * 512 + 64 * 9 * 512 * 512 + 512 * 1024 + 1024 * M,
}


def state_dict(version):
    return random_state_dict(version, M, 1)


dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
crops = torch.from_numpy(rng.integers(0, 256, (N, 80, 80, 1), dtype=np.uint8)).to(dev)
probs = torch.empty((N, M), dtype=torch.float32, device=dev)
stream = torch.cuda.Stream(dev)
for version, precision in (("v118_3", "fp16"), ("v118_3", "bf16x3"), ("v100", "fp16"), ("v110", "fp16"), ("v110", "bf16x3"), ("v100", "fp32"), ("v110", "fp32"), ("v119", "fp32"), ("v200", "fp32")):
    net = trex_b200.VINetwork(M, max_images=N, version=version, precision=precision)
    net.load_weights(state_dict(version))
    with torch.cuda.stream(stream):
        for _ in range(2):
            net.predict_device(crops.data_ptr(), N, 0, probs.data_ptr(), 0, stream.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            net.predict_device(crops.data_ptr(), N, 0, probs.data_ptr(), 0, stream.cuda_stream)
        e1.record(stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"version": version, "precision": precision, "crops": N, "ms": ms, "crops_per_s": N / ms * 1e3,
                      "tflops_algorithmic": 2 * MACS[version] * N / ms / 1e9, "prob_sum": float(probs.sum().item() / N)}))
    net.deinit()
