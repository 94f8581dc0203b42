"""Host mirror of `Python::VINetwork` (Application/src/tracker/ml/VisualIdentification.h:104-189) on top
of the tb_vi_* C ABI: inference only (training stays in the reference).

    VINetwork(num_classes, ...)                       setup(): classes = 0..N-1            (.cpp:98-125)
    load_weights(state_dict)                           load_weights(VIWeights&&)            (.cpp:306-319)
    probabilities(images) -> ndarray (N, M)            probabilities() + transform_results  (.h:115-133, .cpp:809-830)

`state_dict` uses the reference's key names (model.conv1.weight ... model.fc2.bias, torch layouts), i.e.
what `torch.load(..._weights_dict.pth)['state_dict']` holds (T/python/trex_utils.py:65-133).
Calling probabilities() before weights are loaded raises (the reference throws SoftException).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._capi import ViConfig, check, lib

STATE_KEYS = [f"model.conv{i}.{p}" for i in (1, 2, 3) for p in ("weight", "bias")] + \
             [f"model.bn{i}.{p}" for i in (1, 2, 3) for p in ("weight", "bias", "running_mean", "running_var")] + \
             ["model.fc1.weight", "model.fc1.bias", "model.bn4.weight", "model.bn4.bias", "model.fc2.weight", "model.fc2.bias"]

# visual_identification_version -> tb_vi_config.arch (ModelFetcher.add_custom_models, visual_identification_network_torch.py:537-567)
VERSIONS = {"v118_3": 0, "v100": 1, "v110": 2, "v119": 3, "v200": 4}
_BN = ("weight", "bias", "running_mean", "running_var")


def state_keys(version: str):
    """state_dict entries the library needs for a network version (num_batches_tracked is not used in eval mode)."""
    if version == "v118_3":
        return list(STATE_KEYS)
    n_conv, conv_bn, fc_bn = {"v100": (3, False, None), "v110": (3, True, "bn4"), "v119": (4, True, "bn5"), "v200": (5, True, "bn6")}[version]
    keys = [f"model.conv{i}.{p}" for i in range(1, n_conv + 1) for p in ("weight", "bias")]
    if conv_bn:
        keys += [f"model.bn{i}.{p}" for i in range(1, n_conv + 1) for p in _BN]
    if fc_bn:
        keys += [f"model.{fc_bn}.{p}" for p in _BN]
    return keys + ["model.fc1.weight", "model.fc1.bias", "model.fc2.weight", "model.fc2.bias"]


def batch_size_for(num_classes: int) -> int:
    """VisualIdentification.cpp:105-118 (kept for API parity; results do not depend on it in eval mode)."""
    b = max(int(num_classes) if num_classes else 1, 64)
    if b < 128:
        p = 1
        while p < b:
            p <<= 1
        return p
    return 128


class VINetwork:
    def __init__(self, num_classes: int, width=80, height=80, channels=1, max_images=4096, device=0, precision=None, version="v118_3"):
        """version: visual_identification_version ("v118_3" default; "v100" and "v110" share its tensor-core kernels -- v110 only
        for positive BatchNorm scales, so it defaults to fp32 and the tensor path is an opt-in --; "v119", "v200" run in fp32).
        precision (default "fp16c" for v118_3 and v100): "fp16c" (tensor cores, one fp16 MMA + one e5m2 correction MMA per k-step in conv2 /
        conv3: ~2e-5 of fp32 on O(1) logits, within 1e-3 for logits of +-23; default and what bench.py's headline runs), "bf16x3" (3-MMA split,
        ~1e-5: the most accurate), "fp16" (one MMA per k-step, ~3e-4 on O(1) logits: the fastest) or "fp32" (CUDA cores; an independent
        implementation for parity tests)."""
        self.num_classes, self.width, self.height, self.channels = int(num_classes), width, height, channels
        self.max_images = int(max_images)
        if version not in VERSIONS:
            raise ValueError(f"Model {version} not found. Available models are: {list(VERSIONS)}")      # ModelFetcher.get_model
        self.version = version
        if precision is None:
            precision = "fp16c" if version in ("v118_3", "v100") else "fp32"      # v110's tensor path needs positive BatchNorm scales: opt in
        cfg = ViConfig(device=device, width=width, height=height, channels=channels, num_classes=self.num_classes,
                       max_images=self.max_images, precision={"fp32": 0, "bf16x3": 1, "fp16": 2, "fp16c": 3}[precision], arch=VERSIONS[version])
        self._h = C.c_void_p()
        check(lib().tb_vi_create(C.byref(cfg), C.byref(self._h)))
        self.batch_size = batch_size_for(num_classes)

    def load_weights(self, state_dict):
        for k in state_keys(self.version):
            if k not in state_dict:
                raise KeyError(f"state_dict lacks {k}")
            v = state_dict[k]
            a = np.ascontiguousarray(v.detach().cpu().numpy() if hasattr(v, "detach") else v, np.float32)
            check(lib().tb_vi_set_tensor(self._h, k.encode(), a.ctypes.data_as(C.c_void_p), a.size))
        check(lib().tb_vi_commit(self._h))

    def probabilities(self, images, return_logits=False):
        """images: (N,H,W,C) or (N,H,W) uint8 (Image::Ptr list in the reference). Returns softmax rows."""
        x = np.ascontiguousarray(images, np.uint8)
        n = x.shape[0]
        if n == 0:
            z = np.zeros((0, self.num_classes), np.float32)
            return (z, z.copy()) if return_logits else z
        assert x.size == n * self.height * self.width * self.channels, x.shape
        probs = np.empty((n, self.num_classes), np.float32)
        logits = np.empty((n, self.num_classes), np.float32) if return_logits else None
        check(lib().tb_vi_predict(self._h, x.ctypes.data_as(C.c_void_p), n, probs.ctypes.data_as(C.c_void_p),
                                  logits.ctypes.data_as(C.c_void_p) if return_logits else None))
        return (probs, logits) if return_logits else probs

    @staticmethod
    def transform_results(n: int, indexes, values, num_classes: int) -> np.ndarray:
        """VINetwork::transform_results (T/ml/VisualIdentification.cpp:808-828): the flat N x M result of the synchronous
        probabilities() form.  `indexes` lists the images the network returned rows for (ascending), `values[idx]` is the row
        of image idx; images skipped before a listed index get -1 rows, rows after the last listed index stay 0 (as in the
        reference).  With this library every image is classified, so indexes = range(n)."""
        probs = np.zeros((n, num_classes), np.float32)
        i = 0
        for idx in (int(v) for v in indexes):
            if i < idx:
                probs[i:idx] = -1.0
                i = idx
            probs[idx] = np.asarray(values[idx], np.float32)
            i += 1
        return probs

    def paverages(self, ids, images):
        """VINetwork::paverages (T/ml/VisualIdentification.h:146-181): {id: (samples, mean probability row)} over the images of every
        id; rows are added in image order in float32 and divided by float(samples), like the reference's std::transform chain."""
        ids = [int(i) for i in ids]
        probs = self.probabilities(images)
        if len(ids) != len(probs):
            raise ValueError("paverages: ids and images differ in length")
        out = {}
        for i, k in enumerate(ids):
            samples, values = out.get(k, (0, np.zeros(self.num_classes, np.float32)))
            out[k] = (samples + 1, (probs[i] + values).astype(np.float32))
        return {k: (n, (v / np.float32(n)).astype(np.float32)) for k, (n, v) in sorted(out.items())}

    def predict_device(self, images_ptr: int, n_max: int, n_dev_ptr: int, probs_ptr: int, logits_ptr: int = 0, stream: int = 0):
        check(lib().tb_vi_predict_device(self._h, C.c_void_p(images_ptr), n_max, C.c_void_p(n_dev_ptr) if n_dev_ptr else None,
                                         C.c_void_p(probs_ptr), C.c_void_p(logits_ptr) if logits_ptr else None,
                                         C.c_void_p(stream)))

    def set_top1(self, ids_ptr: int, probs_ptr: int):
        """Device buffers (uint32[n], float32[n]) that predict_device fills with the arg-max identity and its probability."""
        check(lib().tb_vi_set_top1(self._h, C.c_void_p(ids_ptr) if ids_ptr else None, C.c_void_p(probs_ptr) if probs_ptr else None))

    def wait(self):
        check(lib().tb_vi_wait(self._h))

    def launch_count(self) -> int:
        return int(lib().tb_vi_launch_count(self._h))

    def profile(self, enable=True):
        check(lib().tb_vi_profile(self._h, int(enable)))

    def kernel_ms(self):
        ms, n = (C.c_double * 5)(), C.c_uint64()
        check(lib().tb_vi_kernel_ms(self._h, C.byref(ms), C.byref(n)))
        return dict(zip(("conv1", "conv2", "conv3", "fc1", "head"), ms)), int(n.value)

    def deinit(self):
        if self._h:
            lib().tb_vi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.deinit()
        except Exception:      # interpreter shutdown
            pass
