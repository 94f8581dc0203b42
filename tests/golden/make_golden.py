#!/usr/bin/env python
"""Generates the committed golden fixtures from the REFERENCE checkout (/root/reference).
Run here (authoring container) only; the GPU box has no reference tree and just reads the .npz.

  testpv_golden.npz  the reference's own segmentation output (videos/test.pv, written by TRex with
                     videos/test.settings: detect_threshold=9, detect_size_filter=[[1,10000]], gray)
                     for two full 2304x2304 frames and eight 512x640 windows, together with the
                     decoded input frames (videos/test_frames/*.jpg) and the background stored in the pv.
  vi_golden.npz      logits / probabilities of the reference's own V118_3 class
                     (visual_identification_network_torch.py, imported from the reference tree)
                     for the state_dict oracle.vi.init_state_dict(seed=0) generates.
  vi_nets_golden.npz the same for the other custom networks (V100, V110, V119, V200).
  posture_golden.npz fish blobs of test.pv with the midline lengths TRex exported for them (compare_data_automatic/*.csv).
  pixels_golden.npz  known-answer vector transcribed from Application/Tests/test_pixels.cpp:1381-1466.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("TREX_REFERENCE", "/root/reference")


def pack(blobs):
    return dict(lines=blobs.lines, pixels=blobs.pixels, line_off=blobs.line_off, px_off=blobs.px_off)


def make_testpv():
    import cv2
    from oracle import seg
    from oracle.pv15 import PV15
    pv = PV15(f"{REF}/videos/test.pv")
    out = {"average": pv.average, "full_frames": np.array([0, 100])}
    for i in (0, 100):
        fr = cv2.imread(f"{REF}/videos/test_frames/frame_{i:03d}.jpg", cv2.IMREAD_UNCHANGED)
        out[f"full{i}_frame"] = fr
        for k, v in pack(pv.frame(i)).items():
            out[f"full{i}_{k}"] = v
    # windows centred on large blobs of other frames
    WH, WW = 512, 640
    wins = []
    for i in (10, 30, 50, 70, 90, 120, 150, 180):
        fr = cv2.imread(f"{REF}/videos/test_frames/frame_{i:03d}.jpg", cv2.IMREAD_UNCHANGED)
        b = pv.frame(i)
        sizes = np.diff(b.px_off)
        k = int(np.argmax(sizes))
        ln, _ = b.blob(k)
        cy, cx = int(ln["y"].mean()), int(ln["x0"].mean())
        y0 = min(max(cy - WH // 2, 0), pv.height - WH); x0 = min(max(cx - WW // 2, 0), pv.width - WW)
        x0 -= x0 % 16
        # expected: pv blobs strictly inside the window (1 px margin) -> window coordinates
        lines, pixels, lo, po = [], [], [0], [0]
        for j in range(len(b)):
            l, p = b.blob(j)
            if l["x0"].min() > x0 and l["x1"].max() < x0 + WW - 1 and l["y"].min() > y0 and l["y"].max() < y0 + WH - 1:
                l = l.copy(); l["x0"] -= x0; l["x1"] -= x0; l["y"] -= y0
                lines.append(l); pixels.append(p); lo.append(lo[-1] + len(l)); po.append(po[-1] + len(p))
        wins.append((i, y0, x0))
        out[f"win{i}_frame"] = fr[y0:y0 + WH, x0:x0 + WW].copy()
        out[f"win{i}_lines"] = np.concatenate(lines); out[f"win{i}_pixels"] = np.concatenate(pixels)
        out[f"win{i}_line_off"] = np.array(lo, np.int64); out[f"win{i}_px_off"] = np.array(po, np.int64)
    out["windows"] = np.array(wins, np.int64)
    np.savez_compressed(os.path.join(HERE, "testpv_golden.npz"), **out)
    print("testpv_golden.npz", os.path.getsize(os.path.join(HERE, "testpv_golden.npz")) // 1024, "KiB")


def make_vi():
    import torch
    sys.dont_write_bytecode = True
    sys.path.insert(0, f"{REF}/Application/src/tracker/python")
    import visual_identification_network_torch as ref_net
    from oracle import vi
    out = {}
    for tag, M, CI in (("m100", 100, 1), ("m8", 8, 1), ("m16c3", 16, 3)):       # c3: rgb8 crops (meta_encoding rgb8)
        sd0 = vi.init_state_dict(M, CI, 80, 80, seed=0, perturb_norm=False)
        torch.manual_seed(0)
        ref = ref_net.ModelFetcher().get_model("v118_3", M, CI, 80, 80, device="cpu")
        ref_sd = ref.state_dict()
        # the oracle's init must consume the RNG exactly like the reference's constructor
        for k, v in sd0.items():
            assert torch.equal(ref_sd[k], v), k
        sd = vi.scale_for_u8_inputs(vi.init_state_dict(M, CI, 80, 80, seed=0, perturb_norm=True))
        ref.load_state_dict(sd); ref.eval()
        rng = np.random.default_rng(7)
        crops = np.zeros((6, 80, 80, CI), np.uint8)
        for n in range(6):     # blob-like: an ellipse of grey values on zero background
            yy, xx = np.mgrid[0:80, 0:80]
            a, b_, th = rng.uniform(12, 30), rng.uniform(4, 9), rng.uniform(0, np.pi)
            u = (xx - 40) * np.cos(th) + (yy - 40) * np.sin(th); v = -(xx - 40) * np.sin(th) + (yy - 40) * np.cos(th)
            m = (u / a) ** 2 + (v / b_) ** 2 <= 1
            for c in range(CI):
                crops[n, ..., c][m] = rng.integers(20, 200, m.sum())
        with torch.no_grad():
            logits = ref(torch.from_numpy(crops).to(torch.float32)).numpy()
            probs = torch.softmax(torch.from_numpy(logits), 1).numpy()
        out[f"{tag}_crops"] = crops; out[f"{tag}_logits"] = logits; out[f"{tag}_probs"] = probs
        out[f"{tag}_checksum"] = np.array(vi.state_checksum(sd))
        mine = vi.forward_logits(sd, crops)
        print(tag, "oracle vs reference class max|dlogit| =", float(np.abs(mine - logits).max()),
              "|logit|max =", float(np.abs(logits).max()))
    np.savez_compressed(os.path.join(HERE, "vi_golden.npz"), **out)


def make_vi_nets():
    """vi_nets_golden.npz: logits / probabilities of the reference's own V100 / V110 / V119 / V200 classes
    (visual_identification_network_torch.py:30-181,262-386) for the state_dicts oracle.vi.init_state_dict_arch generates."""
    import torch
    sys.dont_write_bytecode = True
    sys.path.insert(0, f"{REF}/Application/src/tracker/python")
    import visual_identification_network_torch as ref_net
    from oracle import vi
    out = {}
    for arch in ("v100", "v110", "v119", "v200"):
        for M, CI in ((12, 1), (9, 3)):
            tag = f"{arch}_m{M}c{CI}"
            sd0 = vi.init_state_dict_arch(arch, M, CI, 80, 80, seed=0, perturb_norm=False)
            torch.manual_seed(0)
            ref = ref_net.ModelFetcher().get_model(arch, M, CI, 80, 80, device="cpu")
            ref_sd = ref.state_dict()
            assert set(ref_sd) == set(sd0), (arch, set(ref_sd) ^ set(sd0))
            for k, v in sd0.items():        # same RNG consumption as the reference's constructor
                assert torch.equal(ref_sd[k], v), k
            sd = vi.scale_for_u8_inputs(vi.init_state_dict_arch(arch, M, CI, 80, 80, seed=0))
            ref.load_state_dict(sd); ref.eval()
            rng = np.random.default_rng(11)
            crops = np.zeros((4, 80, 80, CI), np.uint8)
            for n in range(4):
                yy, xx = np.mgrid[0:80, 0:80]
                a, b_, th = rng.uniform(12, 34), rng.uniform(4, 12), rng.uniform(0, np.pi)
                u = (xx - 40) * np.cos(th) + (yy - 40) * np.sin(th); v = -(xx - 40) * np.sin(th) + (yy - 40) * np.cos(th)
                m = (u / a) ** 2 + (v / b_) ** 2 <= 1
                for c in range(CI):
                    crops[n, ..., c][m] = rng.integers(20, 256, m.sum())
            with torch.no_grad():
                logits = ref(torch.from_numpy(crops).to(torch.float32)).numpy()
                probs = torch.softmax(torch.from_numpy(logits), 1).numpy()
            out[f"{tag}_crops"] = crops; out[f"{tag}_logits"] = logits; out[f"{tag}_probs"] = probs
            out[f"{tag}_checksum"] = np.array(vi.state_checksum(sd))
            mine = vi.forward_logits_arch(arch, sd, crops)
            print(tag, "oracle vs reference class max|dlogit| =", float(np.abs(mine - logits).max()), "|logit|max =", float(np.abs(logits).max()))
    np.savez_compressed(os.path.join(HERE, "vi_nets_golden.npz"), **out)
    print("vi_nets_golden.npz", os.path.getsize(os.path.join(HERE, "vi_nets_golden.npz")) // 1024, "KiB")


def make_posture():
    """posture_golden.npz: fish blobs of videos/test.pv (tracker side: re-thresholded with test.settings' track_threshold = 12,
    sign difference, track_size_filter [70, 420)) next to the `midline_length` / `num_pixels` columns TRex itself exported for
    them (videos/compare_data_automatic/test_fish*.csv).  The csv files were written from a slightly different .pv (their pixel
    counts differ by a few pixels from what test.pv yields) and hold the post-processed length rounded to integers, so this is a
    loose, external corroboration of the outline -> midline chain, not a bit-exact pin."""
    import csv
    from oracle import seg
    from oracle.pv15 import PV15
    pv = PV15(f"{REF}/videos/test.pv")
    bg = pv.average
    fish = [list(csv.DictReader(open(f"{REF}/videos/compare_data_automatic/test_fish{k}.csv"))) for k in range(8)]
    dec = lambda b: (b >> 19, (b >> 6) & 0x1FFF)
    out, n = {}, 0
    for f in (0, 1, 2, 50, 100, 150, 199):
        trk = seg.rethreshold(pv.frame(f), bg, 12, method=seg.DIFF_SIGN)
        cand = []
        for k in range(len(trk)):
            l, p = trk.blob(k)
            if 70 <= len(p) < 420:
                cand.append((k, dec(seg.blob_id(l))))
        for k8 in range(8):
            r = fish[k8][f]
            try:
                bid, ml, npx = int(float(r["blobid"])), float(r["midline_length"]), int(float(r["num_pixels"]))
            except (ValueError, OverflowError):      # "inf": the fish was not tracked / had no posture in this frame
                continue
            if not np.isfinite(ml) or ml <= 0:
                continue
            cx, cy = dec(bid)
            best = min(cand, key=lambda c: abs(c[1][0] - cx) + abs(c[1][1] - cy), default=None)
            if best is None or abs(best[1][0] - cx) + abs(best[1][1] - cy) > 6:
                continue
            l, p = trk.blob(best[0])
            if abs(len(p) - npx) > 0.05 * npx:       # the tracker split / merged this blob differently: not the same object
                continue
            x0, y0 = max(int(l["x0"].min()) - 4, 0), max(int(l["y"].min()) - 4, 0)
            x1, y1 = int(l["x1"].max()) + 5, int(l["y"].max()) + 5
            ls = l.copy(); ls["x0"] -= x0; ls["x1"] -= x0; ls["y"] -= y0
            out[f"b{n}_lines"] = ls; out[f"b{n}_pixels"] = p.copy(); out[f"b{n}_bg"] = bg[y0:y1, x0:x1].copy()
            out[f"b{n}_csv"] = np.array([f, k8, ml, npx], np.float64)
            n += 1
    out["count"] = np.array(n)
    np.savez_compressed(os.path.join(HERE, "posture_golden.npz"), **out)
    print("posture_golden.npz", n, "blobs,", os.path.getsize(os.path.join(HERE, "posture_golden.npz")) // 1024, "KiB")


def make_pixels():
    # Application/Tests/test_pixels.cpp:1381-1466 (gray leg): bg gray == the equal-channel BGR values,
    # blob greys = cv::cvtColor of blob_values (only (10,200,10) is not grey: OpenCV fixed point -> 122).
    bg = np.array([[30, 50, 70, 90], [40, 60, 80, 100]], np.uint8)
    px = np.array([25, 110, 80, 122, 30, 95, 200, 100], np.uint8)
    mask = np.array([[0, 255, 0, 255], [0, 255, 255, 0]], np.uint8)
    np.savez(os.path.join(HERE, "pixels_golden.npz"), bg=bg, pixels=px, mask=mask, threshold=25, recount=4)


def make_average():
    """The background stored in videos/test.pv was generated by TRex itself (averaging_method=mode,
    average_samples=100, see the pv's metadata) from videos/test_frames: a 96x128 window of the 100 sampled
    frames (VideoSource.cpp:1040-1060) and of that background pins AveragingAccumulator's mode path."""
    import cv2
    from oracle.pv15 import PV15
    from trex_b200.averaging import sample_indices
    pv = PV15(f"{REF}/videos/test.pv")
    idx = sample_indices(200, 100)
    y0, x0, hh, ww = 1400, 2000, 96, 128
    frames = np.stack([cv2.imread(f"{REF}/videos/test_frames/frame_{i:03d}.jpg", cv2.IMREAD_UNCHANGED)[y0:y0 + hh, x0:x0 + ww] for i in idx])
    np.savez_compressed(os.path.join(HERE, "avg_golden.npz"), frames=frames, expected=pv.average[y0:y0 + hh, x0:x0 + ww],
                        indices=np.array(idx), window=np.array([y0, x0, hh, ww]))
    print("avg_golden.npz", os.path.getsize(os.path.join(HERE, "avg_golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    make_testpv(); make_vi(); make_vi_nets(); make_pixels(); make_average(); make_posture()
