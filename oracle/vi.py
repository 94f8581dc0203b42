"""fp32 CPU restatement of the VisualIdentification inference path.  TEST INFRASTRUCTURE ONLY.

Follows
  V118_3.forward        Application/src/tracker/python/visual_identification_network_torch.py:184-258
  PermuteAxesWrapper    .../visual_identification_network_torch.py:618-644 (NHWC -> NCHW, Normalize = passthrough :7-26)
  predict_numpy         Application/src/tracker/python/visual_recognition_torch.py:290-352
                        (u8 -> float32 with NO scaling :337, model.eval(), softmax(dim=1) :345)
  batch size rule       Application/src/tracker/ml/VisualIdentification.cpp:105-118
  transform_results     .../VisualIdentification.cpp:809-830

The arithmetic itself lives in PyTorch (third party, not vendored in the reference; the reference's
conda env does not pin a version).  Parity is pinned by tests/golden/vi_golden.npz, produced by
tests/golden/make_golden.py from the reference's own V118_3 class (imported from /root/reference)
with the state_dict this module generates; torch fp32 functional ops are used here because this is
a floating-point kernel (conv/linear), tolerance 1e-3 on logits per BASELINE.json north_star.
"""
from __future__ import annotations

import hashlib

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5      # nn.BatchNorm2d default
LN_EPS = 1e-5      # nn.LayerNorm default


def batch_size_for(num_classes: int) -> int:
    """VisualIdentification.cpp:105-118: next_pow2(max(N,64)) if <128 else 128."""
    b = max(int(num_classes) if num_classes else 1, 64)
    if b < 128:
        p = 1
        while p < b:
            p <<= 1
        return p
    return 128


def init_state_dict(num_classes=100, channels=1, width=80, height=80, seed=0, perturb_norm=True):
    """Random-init weights with the SAME RNG consumption order as the reference's constructor
    (V118_3.__init__: conv1, bn1, conv2, bn2, conv3, bn3, fc1, LayerNorm, fc2), so that
    torch.manual_seed(seed) yields the state_dict the reference class gets.  With perturb_norm the
    norm layers' affine/running stats are then drawn from a separate generator so BN folding and
    LayerNorm are really exercised (a fresh BN is the identity)."""
    torch.manual_seed(seed)
    conv1 = torch.nn.Conv2d(channels, 16, 5, padding="same")
    conv2 = torch.nn.Conv2d(16, 64, 5, padding="same")
    conv3 = torch.nn.Conv2d(64, 128, 5, padding="same")
    fc1 = torch.nn.Linear(128 * (width // 8) * (height // 8), 100)
    fc2 = torch.nn.Linear(100, num_classes)
    sd = {}
    for name, m in (("conv1", conv1), ("conv2", conv2), ("conv3", conv3), ("fc1", fc1), ("fc2", fc2)):
        sd[f"model.{name}.weight"] = m.weight.detach().clone()
        sd[f"model.{name}.bias"] = m.bias.detach().clone()
    g = torch.Generator().manual_seed(seed + 12345)
    for name, c in (("bn1", 16), ("bn2", 64), ("bn3", 128)):
        if perturb_norm:
            sd[f"model.{name}.weight"] = 0.5 + torch.rand(c, generator=g)
            sd[f"model.{name}.bias"] = 0.2 * torch.randn(c, generator=g)
            sd[f"model.{name}.running_mean"] = 0.5 * torch.randn(c, generator=g)
            sd[f"model.{name}.running_var"] = 0.5 + torch.rand(c, generator=g)
        else:
            sd[f"model.{name}.weight"] = torch.ones(c); sd[f"model.{name}.bias"] = torch.zeros(c)
            sd[f"model.{name}.running_mean"] = torch.zeros(c); sd[f"model.{name}.running_var"] = torch.ones(c)
        sd[f"model.{name}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    if perturb_norm:
        sd["model.bn4.weight"] = 0.5 + torch.rand(100, generator=g)
        sd["model.bn4.bias"] = 0.2 * torch.randn(100, generator=g)
    else:
        sd["model.bn4.weight"] = torch.ones(100); sd["model.bn4.bias"] = torch.zeros(100)
    return sd


def scale_for_u8_inputs(sd, factor=1.0 / 64.0):
    """Random-init weights fed raw 0..255 inputs saturate nothing but make logits huge; trained
    TRex weights see the same raw range.  Scaling conv1 keeps activations O(1) so that the 1e-3
    logit tolerance is meaningful.  Returns a new dict."""
    out = {k: v.clone() for k, v in sd.items()}
    out["model.conv1.weight"] = out["model.conv1.weight"] * factor
    return out


def state_checksum(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode()); h.update(sd[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def forward_logits(sd, crops_nhwc_u8) -> np.ndarray:
    """crops: (N,H,W,C) uint8 -> logits (N,M) float32."""
    x = torch.from_numpy(np.ascontiguousarray(crops_nhwc_u8)).to(torch.float32)   # no /255 (:337)
    x = x.permute(0, 3, 1, 2).contiguous()
    with torch.no_grad():
        for i in (1, 2, 3):
            x = F.conv2d(x, sd[f"model.conv{i}.weight"], sd[f"model.conv{i}.bias"], padding=2)
            x = F.batch_norm(x, sd[f"model.bn{i}.running_mean"], sd[f"model.bn{i}.running_var"],
                             sd[f"model.bn{i}.weight"], sd[f"model.bn{i}.bias"], training=False, eps=BN_EPS)
            x = F.max_pool2d(F.relu(x), 2)
        x = x.reshape(x.size(0), -1)                                   # NCHW flatten: c*100 + y*10 + x
        x = F.linear(x, sd["model.fc1.weight"], sd["model.fc1.bias"])
        x = F.layer_norm(x, (100,), sd["model.bn4.weight"], sd["model.bn4.bias"], eps=LN_EPS)
        x = F.linear(F.relu(x), sd["model.fc2.weight"], sd["model.fc2.bias"])
    return x.numpy()


def predict(sd, crops_nhwc_u8, batch_size=None) -> np.ndarray:
    """predict_numpy: batched, softmax(dim=1).  Returns probabilities (N,M)."""
    n = len(crops_nhwc_u8)
    m = sd["model.fc2.weight"].shape[0]
    bs = batch_size or batch_size_for(m)
    out = []
    for i in range(0, n, bs):
        lg = torch.from_numpy(forward_logits(sd, crops_nhwc_u8[i:i + bs]))
        out.append(torch.softmax(lg, dim=1).numpy())
    return np.concatenate(out, 0) if out else np.zeros((0, m), np.float32)


def transform_results(n, indexes, values, m):
    """VINetwork::transform_results (.cpp:809-830): rows absent from `indexes` are filled with -1."""
    probs = np.zeros((n, m), np.float32)
    i = 0
    for idx in (int(v) for v in indexes):
        if i < idx:
            probs[i:idx] = -1.0
            i = idx
        probs[idx] = values[idx]
        i += 1
    return probs


# ------------------------------------------------------------------------------------------------
# The other custom networks ModelFetcher offers (visual_identification_network_torch.py:537-567):
#   V200 :30-103, V119 :106-181, V110 :262-325, V100 :328-386.  Same predict path (predict_numpy), eval mode
# (Dropout / Dropout2d are identities).  Each entry lists the conv blocks in constructor order as
# (cout, kernel, pool, batchnorm) -- `bn` "pre" = conv -> BN -> ReLU -> pool (V200 / V119 / V118_3),
# "post" = conv -> pool -> BN -> ReLU (V110), None = conv -> ReLU -> pool (V100) -- then the head.
# ------------------------------------------------------------------------------------------------
ARCHS = {
    "v100": dict(convs=[(16, 5, 2, None), (64, 5, 2, None), (100, 5, 2, None)], gap=False, fc1=100, fc_bn=None),
    "v110": dict(convs=[(16, 5, 2, "post"), (64, 5, 2, "post"), (100, 5, 2, "post")], gap=False, fc1=100, fc_bn="bn4"),
    "v119": dict(convs=[(256, 5, 2, "pre"), (128, 5, 2, "pre"), (32, 5, 2, "pre"), (128, 5, 2, "pre")], gap=False, fc1=1024, fc_bn="bn5"),
    "v200": dict(convs=[(64, 3, 1, "pre"), (128, 3, 3, "pre"), (256, 3, 1, "pre"), (512, 3, 3, "pre"), (512, 3, 3, "pre")],
                 gap=True, fc1=1024, fc_bn="bn6"),
}


def arch_fc1_in(arch, width=80, height=80):
    a = ARCHS[arch]
    w, h = width, height
    for _, _, pool, _ in a["convs"]:
        w, h = w // pool, h // pool
    c = a["convs"][-1][0]
    return c if a["gap"] else c * w * h


def init_state_dict_arch(arch, num_classes=100, channels=1, width=80, height=80, seed=0, perturb_norm=True):
    """Random-init state_dict of V100 / V110 / V119 / V200 with the RNG consumption order of the reference's constructors
    (parameters are drawn layer by layer in __init__ order: the convs, fc1, fc2; norm layers draw nothing)."""
    a = ARCHS[arch]
    torch.manual_seed(seed)
    sd, cin, mods = {}, channels, []
    for i, (cout, ks, _, _) in enumerate(a["convs"], 1):
        mods.append((f"conv{i}", torch.nn.Conv2d(cin, cout, ks, padding="same")))
        cin = cout
    mods.append(("fc1", torch.nn.Linear(arch_fc1_in(arch, width, height), a["fc1"])))
    mods.append(("fc2", torch.nn.Linear(a["fc1"], num_classes)))
    for name, m in mods:
        sd[f"model.{name}.weight"] = m.weight.detach().clone()
        sd[f"model.{name}.bias"] = m.bias.detach().clone()
    g = torch.Generator().manual_seed(seed + 12345)
    norms = [(f"bn{i}", cout) for i, (cout, _, _, bn) in enumerate(a["convs"], 1) if bn]
    if a["fc_bn"]:
        norms.append((a["fc_bn"], a["fc1"]))
    for name, c in norms:
        if perturb_norm:
            sd[f"model.{name}.weight"] = 0.5 + torch.rand(c, generator=g)
            sd[f"model.{name}.bias"] = 0.2 * torch.randn(c, generator=g)
            sd[f"model.{name}.running_mean"] = 0.5 * torch.randn(c, generator=g)
            sd[f"model.{name}.running_var"] = 0.5 + torch.rand(c, generator=g)
        else:
            sd[f"model.{name}.weight"] = torch.ones(c); sd[f"model.{name}.bias"] = torch.zeros(c)
            sd[f"model.{name}.running_mean"] = torch.zeros(c); sd[f"model.{name}.running_var"] = torch.ones(c)
        sd[f"model.{name}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    return sd


def forward_logits_arch(arch, sd, crops_nhwc_u8) -> np.ndarray:
    """forward() of V100 / V110 / V119 / V200 behind PermuteAxesWrapper, eval mode: (N,H,W,C) uint8 -> logits (N,M)."""
    a = ARCHS[arch]
    x = torch.from_numpy(np.ascontiguousarray(crops_nhwc_u8)).to(torch.float32).permute(0, 3, 1, 2).contiguous()

    def bn(x, name):
        return F.batch_norm(x, sd[f"model.{name}.running_mean"], sd[f"model.{name}.running_var"],
                            sd[f"model.{name}.weight"], sd[f"model.{name}.bias"], training=False, eps=BN_EPS)
    with torch.no_grad():
        for i, (_, ks, pool, mode) in enumerate(a["convs"], 1):
            x = F.conv2d(x, sd[f"model.conv{i}.weight"], sd[f"model.conv{i}.bias"], padding=ks // 2)
            if mode == "pre":
                x = F.relu(bn(x, f"bn{i}"))
                if pool > 1:
                    x = F.max_pool2d(x, pool)
            elif mode == "post":
                x = F.relu(bn(F.max_pool2d(x, pool), f"bn{i}"))
            else:
                x = F.max_pool2d(F.relu(x), pool)
        if a["gap"]:
            x = F.adaptive_avg_pool2d(x, (1, 1))
        x = x.reshape(x.size(0), -1)
        x = F.linear(x, sd["model.fc1.weight"], sd["model.fc1.bias"])
        if a["fc_bn"]:
            x = bn(x, a["fc_bn"])
        x = F.linear(F.relu(x), sd["model.fc2.weight"], sd["model.fc2.bias"])
    return x.numpy()


def predict_arch(arch, sd, crops_nhwc_u8) -> np.ndarray:
    lg = torch.from_numpy(forward_logits_arch(arch, sd, crops_nhwc_u8))
    return torch.softmax(lg, dim=1).numpy()
