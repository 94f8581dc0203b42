"""GPU parity tests of the identification path: CUDA V118_3 vs the fp32 oracle and the outputs of the
reference's own class (tests/golden/vi_golden.npz).  Tolerance: 1e-3 on logits and probabilities
(BASELINE.json north_star)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
TOL = 1e-3


def logit_tol(precision, ref):
    """ABSOLUTE 1e-3 on logits (BASELINE.json north_star) for every precision that is meant to hold it at any logit scale; the single-MMA
    "fp16" mode carries a relative error of ~3e-4 of the logit scale and is held to 1e-3 * max(1, max|logit|) (tests/test_gpu_chain.py)."""
    return TOL * max(1.0, float(np.abs(ref).max())) if precision == "fp16" else TOL


PRECISIONS = ["fp32", "bf16x3", "fp16", "fp16c"]


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("tag,M", [("m100", 100), ("m8", 8)])
def test_reference_class_golden(tag, M, precision):
    import trex_b200
    from oracle import vi
    g = np.load(os.path.join(GOLDEN, "vi_golden.npz"))
    sd = vi.scale_for_u8_inputs(vi.init_state_dict(M, 1, 80, 80, seed=0))
    assert vi.state_checksum(sd) == str(g[f"{tag}_checksum"])
    net = trex_b200.VINetwork(M, max_images=16, precision=precision)
    net.load_weights(sd)
    probs, logits = net.probabilities(g[f"{tag}_crops"], return_logits=True)
    assert np.abs(logits - g[f"{tag}_logits"]).max() < TOL
    assert np.abs(probs - g[f"{tag}_probs"]).max() < TOL
    assert np.allclose(probs.sum(1), 1, atol=1e-5)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_batch_vs_oracle_and_chunking(precision):
    import trex_b200
    from oracle import vi
    M = 100
    sd = vi.scale_for_u8_inputs(vi.init_state_dict(M, 1, 80, 80, seed=0))
    rng = np.random.default_rng(5)
    crops = np.zeros((70, 80, 80, 1), np.uint8)
    for n in range(70):
        h, w = rng.integers(8, 70), rng.integers(8, 70)
        y, x = rng.integers(0, 80 - h), rng.integers(0, 80 - w)
        crops[n, y:y + h, x:x + w, 0] = rng.integers(0, 256, (h, w))
    crops[0] = 0; crops[1] = 255
    net = trex_b200.VINetwork(M, max_images=32, precision=precision)       # forces 3 chunks through the C ABI
    net.load_weights(sd)
    probs, logits = net.probabilities(crops, return_logits=True)
    ref = vi.forward_logits(sd, crops)
    scale = max(1.0, float(np.abs(ref).max()))
    assert np.abs(logits - ref).max() < TOL * scale
    assert np.abs(probs - vi.predict(sd, crops)).max() < TOL
    assert (probs.argmax(1) == ref.argmax(1)).mean() > 0.98


def test_errors():
    import trex_b200
    net = trex_b200.VINetwork(10, max_images=4)
    with pytest.raises(trex_b200.TrexB200Error) as e:
        net.probabilities(np.zeros((1, 80, 80, 1), np.uint8))
    assert e.value.code == -3                            # no weights: reference throws SoftException
    assert net.probabilities(np.zeros((0, 80, 80, 1), np.uint8)).shape == (0, 10)
    with pytest.raises(trex_b200.TrexB200Error):
        trex_b200.VINetwork(10, width=64, height=64)


def test_tensor_path_many_images_persistent_ctas():
    """More images than SMs: persistent CTAs loop, weight/accumulator barriers wrap their phases."""
    import trex_b200
    from oracle import vi
    M = 100
    sd = vi.scale_for_u8_inputs(vi.init_state_dict(M, 1, 80, 80, seed=0))
    rng = np.random.default_rng(17)
    n = 333
    crops = np.zeros((n, 80, 80, 1), np.uint8)
    for i in range(n):
        h, w = rng.integers(10, 60), rng.integers(10, 60)
        y, x = rng.integers(0, 80 - h), rng.integers(0, 80 - w)
        crops[i, y:y + h, x:x + w, 0] = rng.integers(1, 256, (h, w))
    net = trex_b200.VINetwork(M, max_images=512, precision="bf16x3")
    net.load_weights(sd)
    probs, logits = net.probabilities(crops, return_logits=True)
    ref = vi.forward_logits(sd, crops)
    assert np.abs(logits - ref).max() < logit_tol("bf16x3", ref)
    assert np.abs(probs - vi.predict(sd, crops)).max() < TOL
    probs2 = net.probabilities(crops[:100])          # handle reuse, different n
    assert np.abs(probs2 - probs[:100]).max() < 1e-6


@pytest.mark.parametrize("precision", PRECISIONS)
def test_256_classes(precision):
    """BASELINE configs[3]: 256 individuals -> M = 256 classes."""
    import trex_b200
    from oracle import vi
    M = 256
    sd = vi.scale_for_u8_inputs(vi.init_state_dict(M, 1, 80, 80, seed=3))
    rng = np.random.default_rng(2)
    crops = np.zeros((40, 80, 80, 1), np.uint8)
    for i in range(40):
        crops[i, 20:60, 10 + i % 20:50 + i % 20, 0] = rng.integers(1, 200, (40, 40))
    net = trex_b200.VINetwork(M, max_images=64, precision=precision)
    net.load_weights(sd)
    probs, logits = net.probabilities(crops, return_logits=True)
    ref = vi.forward_logits(sd, crops)
    assert np.abs(logits - ref).max() < logit_tol(precision, ref)
    assert np.abs(probs - vi.predict(sd, crops)).max() < TOL


def test_top1_device_outputs():
    """Arg-max identity + its probability written on the device (metadata for Tracker::predicted / the all-gather)."""
    import torch
    import trex_b200
    from oracle import vi
    M = 100
    sd = vi.scale_for_u8_inputs(vi.init_state_dict(M, 1, 80, 80, seed=0))
    rng = np.random.default_rng(9)
    crops = np.zeros((50, 80, 80, 1), np.uint8)
    for i in range(50):
        crops[i, 10:70, 20:60, 0] = rng.integers(1, 255, (60, 40))
    dev = torch.device("cuda", 0)
    x = torch.from_numpy(crops).to(dev)
    probs = torch.empty((50, M), dtype=torch.float32, device=dev)
    ids = torch.zeros(50, dtype=torch.int32, device=dev); p = torch.zeros(50, dtype=torch.float32, device=dev)
    net = trex_b200.VINetwork(M, max_images=64, precision="bf16x3")
    net.load_weights(sd)
    net.set_top1(ids.data_ptr(), p.data_ptr())
    s = torch.cuda.Stream(dev)
    net.predict_device(x.data_ptr(), 50, 0, probs.data_ptr(), 0, s.cuda_stream)
    net.wait()
    pr = probs.cpu().numpy()
    assert np.array_equal(ids.cpu().numpy(), pr.argmax(1))
    assert np.allclose(p.cpu().numpy(), pr.max(1), rtol=1e-6, atol=1e-7)
    assert np.abs(pr - vi.predict(sd, crops)).max() < TOL


@pytest.mark.parametrize("precision", PRECISIONS)
def test_rgb8_crops_three_input_channels(precision):
    """meta_encoding rgb8: 80x80x3 crops (NHWC), conv1 3->16.  Golden = the reference's own V118_3(channels=3)."""
    import trex_b200
    from oracle import vi
    g = np.load(os.path.join(GOLDEN, "vi_golden.npz"))
    M = 16
    sd = vi.scale_for_u8_inputs(vi.init_state_dict(M, 3, 80, 80, seed=0))
    assert vi.state_checksum(sd) == str(g["m16c3_checksum"])
    net = trex_b200.VINetwork(M, channels=3, max_images=64, precision=precision)
    net.load_weights(sd)
    probs, logits = net.probabilities(g["m16c3_crops"], return_logits=True)
    assert np.abs(logits - g["m16c3_logits"]).max() < TOL
    assert np.abs(probs - g["m16c3_probs"]).max() < TOL
    rng = np.random.default_rng(23)
    n = 200                                              # more crops than SMs, 4 chunks
    crops = np.zeros((n, 80, 80, 3), np.uint8)
    for i in range(n):
        h, w = rng.integers(8, 70), rng.integers(8, 70)
        y, x = rng.integers(0, 80 - h), rng.integers(0, 80 - w)
        crops[i, y:y + h, x:x + w] = rng.integers(0, 256, (h, w, 3))
    crops[0] = 0; crops[1] = 255; crops[2, ..., 0] = 255; crops[3, ..., 2] = 255
    probs, logits = net.probabilities(crops, return_logits=True)
    ref = vi.forward_logits(sd, crops)
    assert np.abs(logits - ref).max() < logit_tol(precision, ref)
    assert np.abs(probs - vi.predict(sd, crops)).max() < TOL


def test_precision_margins():
    """Measured max|dlogit| of the tensor-core precisions against the fp32 oracle on blob-like crops (tolerance 1e-3):
    bf16x3 (three MMAs per k-step) ~1e-5, fp16 (one MMA per k-step in conv2 / conv3) a few 1e-4, fp16c (fp16 + one e5m2
    correction MMA per k-step) in between."""
    import trex_b200
    from oracle import vi
    M = 100
    rng = np.random.default_rng(99)
    n = 256
    crops = np.zeros((n, 80, 80, 1), np.uint8)
    yy, xx = np.mgrid[0:80, 0:80]
    for i in range(n):
        a, b_, th = rng.uniform(12, 30), rng.uniform(4, 9), rng.uniform(0, np.pi)
        u = (xx - 40) * np.cos(th) + (yy - 40) * np.sin(th); v = -(xx - 40) * np.sin(th) + (yy - 40) * np.cos(th)
        m = (u / a) ** 2 + (v / b_) ** 2 <= 1
        crops[i, ..., 0][m] = rng.integers(20, 200, int(m.sum()))
    worst = {}
    for seed in (0, 1, 2):
        sd = vi.scale_for_u8_inputs(vi.init_state_dict(M, 1, 80, 80, seed=seed))
        ref = vi.forward_logits(sd, crops)
        for precision in ("bf16x3", "fp16", "fp16c"):
            net = trex_b200.VINetwork(M, max_images=256, precision=precision)
            net.load_weights(sd)
            _, logits = net.probabilities(crops, return_logits=True)
            worst[precision] = max(worst.get(precision, 0.0), float(np.abs(logits - ref).max()))
            net.deinit()
    print("max|dlogit|", worst)
    assert worst["bf16x3"] < 1e-4
    assert worst["fp16"] < TOL
    assert worst["fp16c"] < 0.25 * worst["fp16"] and worst["fp16c"] < 1.5e-4     # fp16 + e5m2 corrections: ~2^-3 of fp16's rounding error


@pytest.mark.parametrize("variant", ["0", "1"])
def test_conv2_fallback_kernels(variant):
    """TB_VI_CONV2_PAIR selects conv2's kernel for fp16 / fp16c (2, the default everywhere else: tap pairs on CTA pairs; 1: tap pairs on single
    CTAs; 0: one tap per MMA).  The switch is read once per process, so the two fall-backs run in a child process: same golden logits."""
    import subprocess
    import sys
    code = r"""
import os, sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import trex_b200
from oracle import vi
g = np.load(%r)
sd = vi.scale_for_u8_inputs(vi.init_state_dict(100, 1, 80, 80, seed=0))
for precision, tol in (("fp16c", 1e-3), ("fp16", 1e-3)):
    net = trex_b200.VINetwork(100, max_images=16, precision=precision)
    net.load_weights(sd)
    for n in (6, 3):
        probs, logits = net.probabilities(g["m100_crops"][:n], return_logits=True)
        d = float(np.abs(logits - g["m100_logits"][:n]).max())
        assert d < tol, (precision, n, d)
print("ok")
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)), os.path.join(GOLDEN, "vi_golden.npz"))
    env = dict(os.environ, TB_VI_CONV2_PAIR=variant)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("precision", ["fp16", "fp16c"])
def test_odd_crop_counts_cta_pair_dummy_item(precision):
    """conv2 on CTA pairs walks 3 bands per crop two at a time: an odd number of crops leaves one CTA of the last pair with a dummy item (it must
    neither store nor stall its peer).  1, 3 and 5 crops -> 3, 9, 15 items; results equal the same crops inside an even batch."""
    import trex_b200
    from oracle import vi
    g = np.load(os.path.join(GOLDEN, "vi_golden.npz"))
    sd = vi.scale_for_u8_inputs(vi.init_state_dict(100, 1, 80, 80, seed=0))
    net = trex_b200.VINetwork(100, max_images=16, precision=precision)
    net.load_weights(sd)
    _, full = net.probabilities(g["m100_crops"], return_logits=True)
    assert np.abs(full - g["m100_logits"]).max() < TOL
    for n in (1, 3, 5):
        _, logits = net.probabilities(g["m100_crops"][:n], return_logits=True)
        assert np.array_equal(logits, full[:n]), n
