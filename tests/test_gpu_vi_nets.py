"""GPU parity tests of the other custom identification networks (visual_identification_version v100 / v110 / v119 / v200,
T/python/visual_identification_network_torch.py:30-181,262-386) against the outputs of the reference's own classes
(tests/golden/vi_nets_golden.npz) and the fp32 oracle.  Tolerance: 1e-3 on logits and probabilities (BASELINE.json north_star)."""
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
ARCHS = ["v100", "v110", "v119", "v200"]


@pytest.mark.parametrize("M,CI", [(12, 1), (9, 3)])
@pytest.mark.parametrize("arch", ARCHS)
def test_reference_class_golden(arch, M, CI):
    import trex_b200
    from oracle import vi
    g = np.load(os.path.join(GOLDEN, "vi_nets_golden.npz"))
    tag = f"{arch}_m{M}c{CI}"
    sd = vi.scale_for_u8_inputs(vi.init_state_dict_arch(arch, M, CI, 80, 80, seed=0))
    assert vi.state_checksum(sd) == str(g[f"{tag}_checksum"])
    net = trex_b200.VINetwork(M, channels=CI, max_images=8, version=arch, precision="fp32")
    net.load_weights(sd)
    probs, logits = net.probabilities(g[f"{tag}_crops"], return_logits=True)
    assert np.abs(logits - g[f"{tag}_logits"]).max() < TOL
    assert np.abs(probs - g[f"{tag}_probs"]).max() < TOL
    assert np.allclose(probs.sum(1), 1, atol=1e-5)


@pytest.mark.parametrize("arch", ARCHS)
def test_batch_vs_oracle_and_chunking(arch):
    """More images than one predict call holds (max_images) and than one CTA group (images per block) covers; 100 classes."""
    import trex_b200
    from oracle import vi
    M = 100
    sd = vi.scale_for_u8_inputs(vi.init_state_dict_arch(arch, M, 1, 80, 80, seed=3))
    rng = np.random.default_rng(5)
    n = 37
    crops = np.zeros((n, 80, 80, 1), np.uint8)
    for i in range(n):
        h, w = rng.integers(8, 70), rng.integers(8, 70)
        y, x = rng.integers(0, 80 - h), rng.integers(0, 80 - w)
        crops[i, y:y + h, x:x + w, 0] = rng.integers(0, 256, (h, w))
    crops[0] = 0; crops[1] = 255
    net = trex_b200.VINetwork(M, max_images=20, version=arch, precision="fp32")
    net.load_weights(sd)
    probs, logits = net.probabilities(crops, return_logits=True)
    ref = vi.forward_logits_arch(arch, sd, crops)
    scale = max(1.0, float(np.abs(ref).max()))
    assert np.abs(logits - ref).max() < TOL * scale
    assert np.abs(probs - vi.predict_arch(arch, sd, crops)).max() < TOL
    p2 = net.probabilities(crops[:5])                        # probabilities alone (no logits buffer)
    assert np.array_equal(p2, probs[:5])


def test_version_errors():
    import trex_b200
    with pytest.raises(ValueError):
        trex_b200.VINetwork(10, version="v999")
    with pytest.raises(trex_b200.TrexB200Error):
        trex_b200.VINetwork(10, version="v119", precision="fp16")       # tensor-core precisions: v118_3, v100, v110
    net = trex_b200.VINetwork(10, max_images=4, version="v110")
    with pytest.raises(trex_b200.TrexB200Error) as e:
        net.probabilities(np.zeros((1, 80, 80, 1), np.uint8))
    assert e.value.code == -3                                             # no weights loaded
    from oracle import vi
    sd = vi.init_state_dict_arch("v110", 10, 1)
    del sd["model.bn4.running_var"]
    with pytest.raises(KeyError):
        net.load_weights(sd)


@pytest.mark.parametrize("precision", ["bf16x3", "fp16"])
@pytest.mark.parametrize("M,CI", [(12, 1), (9, 3)])
@pytest.mark.parametrize("arch", ["v100", "v110"])
def test_v100_v110_on_the_tensor_path(arch, M, CI, precision):
    """v100 / v110 have v118_3's layer shapes up to conv3's width (100 channels, zero-padded to 128) and the head's norm, so
    they run on the same tcgen05 kernels; checked against the outputs of the reference's own classes."""
    import trex_b200
    from oracle import vi
    g = np.load(os.path.join(GOLDEN, "vi_nets_golden.npz"))
    tag = f"{arch}_m{M}c{CI}"
    sd = vi.scale_for_u8_inputs(vi.init_state_dict_arch(arch, M, CI, 80, 80, seed=0))
    net = trex_b200.VINetwork(M, channels=CI, max_images=8, version=arch, precision=precision)
    net.load_weights(sd)
    probs, logits = net.probabilities(g[f"{tag}_crops"], return_logits=True)
    assert np.abs(logits - g[f"{tag}_logits"]).max() < TOL
    assert np.abs(probs - g[f"{tag}_probs"]).max() < TOL


@pytest.mark.parametrize("arch", ["v100", "v110"])
def test_tensor_path_batch_vs_oracle(arch):
    import trex_b200
    from oracle import vi
    M = 100
    sd = vi.scale_for_u8_inputs(vi.init_state_dict_arch(arch, M, 1, 80, 80, seed=3))
    rng = np.random.default_rng(8)
    n = 200
    crops = np.zeros((n, 80, 80, 1), np.uint8)
    for i in range(n):
        h, w = rng.integers(8, 70), rng.integers(8, 70)
        y, x = rng.integers(0, 80 - h), rng.integers(0, 80 - w)
        crops[i, y:y + h, x:x + w, 0] = rng.integers(0, 256, (h, w))
    ref = vi.forward_logits_arch(arch, sd, crops)
    for precision in ("bf16x3", "fp16"):
        net = trex_b200.VINetwork(M, max_images=256, version=arch, precision=precision)
        net.load_weights(sd)
        probs, logits = net.probabilities(crops, return_logits=True)
        assert np.abs(logits - ref).max() < logit_tol(precision, ref), precision
        assert np.abs(probs - vi.predict_arch(arch, sd, crops)).max() < TOL


def test_v110_tensor_path_refuses_negative_batchnorm_scales():
    """BatchNorm follows the max-pool in V110: folding it into the filters is only valid for positive scales."""
    import trex_b200
    from oracle import vi
    sd = vi.init_state_dict_arch("v110", 10, 1, seed=0)
    sd["model.bn2.weight"] = sd["model.bn2.weight"].clone(); sd["model.bn2.weight"][5] = -0.3
    net = trex_b200.VINetwork(10, max_images=4, version="v110", precision="bf16x3")
    with pytest.raises(trex_b200.TrexB200Error):
        net.load_weights(sd)
    crops = np.random.default_rng(0).integers(0, 256, (3, 80, 80, 1), dtype=np.uint8)
    net32 = trex_b200.VINetwork(10, max_images=4, version="v110")          # fp32: any sign
    net32.load_weights(sd)
    assert np.abs(net32.probabilities(crops) - vi.predict_arch("v110", sd, crops)).max() < TOL
