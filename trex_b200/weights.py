"""Weight plumbing for VINetwork: random initialisation with torch's default layer init (what the
reference gets from constructing V118_3, visual_identification_network_torch.py:189-214) and loading of
the reference's checkpoints (`*_weights_dict.pth` = {"state_dict", "metadata"} or a bare state_dict,
T/python/trex_utils.py:65-133).  Host-side only; no compute on the path."""
from __future__ import annotations

import torch


def random_v118_3_state_dict(num_classes=100, channels=1, width=80, height=80, seed=0, input_scale=1.0 / 64.0):
    """state_dict with the reference's key names.  conv1 is scaled by `input_scale` so that raw 0..255
    inputs give O(1) activations, as trained weights would."""
    torch.manual_seed(seed)
    layers = {
        "conv1": torch.nn.Conv2d(channels, 16, 5, padding="same"),
        "conv2": torch.nn.Conv2d(16, 64, 5, padding="same"),
        "conv3": torch.nn.Conv2d(64, 128, 5, padding="same"),
        "fc1": torch.nn.Linear(128 * (width // 8) * (height // 8), 100),
        "fc2": torch.nn.Linear(100, num_classes),
    }
    sd = {}
    for name, m in layers.items():
        sd[f"model.{name}.weight"] = m.weight.detach().clone()
        sd[f"model.{name}.bias"] = m.bias.detach().clone()
    sd["model.conv1.weight"] *= input_scale
    g = torch.Generator().manual_seed(seed + 12345)
    for name, c in (("bn1", 16), ("bn2", 64), ("bn3", 128)):
        sd[f"model.{name}.weight"] = 0.5 + torch.rand(c, generator=g)
        sd[f"model.{name}.bias"] = 0.2 * torch.randn(c, generator=g)
        sd[f"model.{name}.running_mean"] = 0.5 * torch.randn(c, generator=g)
        sd[f"model.{name}.running_var"] = 0.5 + torch.rand(c, generator=g)
    sd["model.bn4.weight"] = 0.5 + torch.rand(100, generator=g)
    sd["model.bn4.bias"] = 0.2 * torch.randn(100, generator=g)
    return sd


# conv blocks (cout, kernel, pool, has BatchNorm2d), fc1 width, BatchNorm1d after fc1, global average pool -- the other custom
# networks of ModelFetcher (visual_identification_network_torch.py:30-181,262-386)
_ARCHS = {
    "v100": ([(16, 5, 2, False), (64, 5, 2, False), (100, 5, 2, False)], 100, None, False),
    "v110": ([(16, 5, 2, True), (64, 5, 2, True), (100, 5, 2, True)], 100, "bn4", False),
    "v119": ([(256, 5, 2, True), (128, 5, 2, True), (32, 5, 2, True), (128, 5, 2, True)], 1024, "bn5", False),
    "v200": ([(64, 3, 1, True), (128, 3, 3, True), (256, 3, 1, True), (512, 3, 3, True), (512, 3, 3, True)], 1024, "bn6", True),
}


def random_state_dict(version, num_classes=100, channels=1, width=80, height=80, seed=0, input_scale=1.0 / 64.0):
    """Random-init state_dict (torch's default layer init, perturbed norm layers) for any supported visual_identification_version."""
    if version == "v118_3":
        return random_v118_3_state_dict(num_classes, channels, width, height, seed, input_scale)
    convs, fc1_out, fc_bn, gap = _ARCHS[version]
    torch.manual_seed(seed)
    g = torch.Generator().manual_seed(seed + 12345)
    sd, cin, w, h = {}, channels, width, height

    def norm(name, c):
        sd[f"model.{name}.weight"] = 0.5 + torch.rand(c, generator=g)
        sd[f"model.{name}.bias"] = 0.2 * torch.randn(c, generator=g)
        sd[f"model.{name}.running_mean"] = 0.5 * torch.randn(c, generator=g)
        sd[f"model.{name}.running_var"] = 0.5 + torch.rand(c, generator=g)
    for i, (cout, ks, pool, bn) in enumerate(convs, 1):
        m = torch.nn.Conv2d(cin, cout, ks, padding="same")
        sd[f"model.conv{i}.weight"], sd[f"model.conv{i}.bias"] = m.weight.detach().clone(), m.bias.detach().clone()
        if bn:
            norm(f"bn{i}", cout)
        cin, w, h = cout, w // pool, h // pool
    for name, m in (("fc1", torch.nn.Linear(cin if gap else cin * w * h, fc1_out)), ("fc2", torch.nn.Linear(fc1_out, num_classes))):
        sd[f"model.{name}.weight"], sd[f"model.{name}.bias"] = m.weight.detach().clone(), m.bias.detach().clone()
    if fc_bn:
        norm(fc_bn, fc1_out)
    sd["model.conv1.weight"] *= input_scale
    return sd


def load_reference_checkpoint(path):
    """Reads a TRex VI checkpoint and returns its state_dict (keys model.*)."""
    obj = torch.load(path, map_location="cpu", weights_only=True)
    sd = obj["state_dict"] if isinstance(obj, dict) and "state_dict" in obj else obj
    return {k: v for k, v in sd.items() if not k.endswith("num_batches_tracked")}
