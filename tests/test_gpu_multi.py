"""The multi-GPU data plane on hardware (needs >= 2 GPUs; skipped on a single-GPU box): torchrun with two ranks, NCCL all-gather of
the in-place metadata blocks, every rank checks every other rank's block and the unpacked frames (tests/multi_gpu_worker.py)."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_gpu_gathered_metadata_matches_every_rank():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_GPU_META_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
