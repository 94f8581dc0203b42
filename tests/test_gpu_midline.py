"""GPU parity of the midline stage (N4, second stage): tb_seg_midlines vs the oracle's restatement of Outline::smooth,
offset_to_middle and calculate_midline (T/tracking/Outline.cpp:330-452,454-718,768-868; C/misc/CircularGraph.cpp:12-606).
Bit-exact: both sides evaluate the reference's float code without contraction."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(bs, frames, rd, every=1, **kw):
    from oracle import posture, seg
    got = bs.apply(frames)
    mids = bs.midlines(rd, **kw)
    n = sum(len(g) for g in got)
    assert len(mids) == n and n > 0
    P = posture.default_params(**kw)
    k = ok = 0
    for blobs in got:
        for b in blobs:
            if k % every == 0:
                segs, tail, head, pts = mids[k]
                ol = seg.outline_resample(seg.longest_outline(b.lines), rd)
                try:
                    rs, rt, rh, rp = posture.calculate_midline(ol, P)
                except ValueError:
                    assert len(segs) == 0, k
                else:
                    assert (tail, head) == (rt, rh), k
                    assert np.array_equal(pts, rp), k
                    assert np.array_equal(segs, rs), k
                    ok += 1
            k += 1
    return n, ok


def test_midlines_of_the_benchmark_workload():
    """1080p, 100 elongated individuals per frame: every blob's midline equals the oracle's, bit for bit."""
    import trex_b200
    from trex_b200.synthetic import BlobWorld
    world = BlobWorld(n_blobs=100, seed=6)
    frames = world.frames(2)
    bs = trex_b200.BackgroundSubtraction(world.bg, settings=trex_b200.DetectSettings(), max_batch=2)
    n, ok = _check(bs, frames, 1.0)
    assert n > 150 and ok > 0.9 * n
    # an elongated ellipse: tail and head sit at opposite ends of the outline
    segs, tail, head, pts = bs.midlines(1.0)[0]
    assert tail == 0 and abs(head - len(pts) / 2) < len(pts) * 0.2 and len(segs) > 10


@pytest.mark.parametrize("rd,kw", [(0.5, {}), (1.0, dict(outline_approximate=0)), (1.0, dict(outline_smooth_samples=0, midline_invert=1)),
                                   (2.0, dict(outline_smooth_samples=2, outline_smooth_step=2, midline_start_with_head=1))])
def test_midlines_of_noisy_blobs_and_settings(rd, kw):
    """Ragged blobs (holes, single pixels, tiny outlines -> 'too few segments') under non-default settings."""
    import trex_b200
    rng = np.random.default_rng(int(rd * 10) + len(kw))
    h, w = 96, 160
    bg = np.zeros((h, w), np.uint8)
    frames = [np.where(rng.random((h, w)) < d, 200, 0).astype(np.uint8) for d in (0.3, 0.55)]
    s = trex_b200.DetectSettings(detect_threshold=15, detect_size_filter=[])
    bs = trex_b200.BackgroundSubtraction(bg, settings=s, max_batch=2, max_runs_per_frame=h * w // 2 + 16, max_pixels_per_frame=h * w)
    n, ok = _check(bs, frames, rd, **kw)
    assert n > 200 and ok > 20


def test_midline_errors():
    import trex_b200
    bg = np.full((64, 64), 100, np.uint8)
    bs = trex_b200.BackgroundSubtraction(bg, max_batch=1)
    fr = bg.copy(); fr[20:30, 10:50] = 20
    bs.apply([fr])
    with pytest.raises(trex_b200.TrexB200Error):
        bs.midlines(1.0, peak_mode=1)                       # broad tails are not built
    assert len(bs.midlines(1.0)) == 1


def test_midline_lengths_against_the_references_own_export_gpu():
    """GPU leg of tests/test_oracle_posture.py::test_midline_lengths_against_the_references_own_export: the golden fish blobs of
    videos/test.pv are pasted into frames over their background windows, segmented with the sign difference at threshold 8
    (> 8 == >= 9 = track_posture_threshold), and the raw midline length from tb_seg_midlines is compared with TRex's exported
    midline_length."""
    import os
    import trex_b200
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "posture_golden.npz"))
    n = int(g["count"])
    H, W = 96, 128
    ratios = []
    for i in range(n):
        lines, px, bgw = g[f"b{i}_lines"], g[f"b{i}_pixels"], g[f"b{i}_bg"]
        bg = np.full((H, W), 200, np.uint8); bg[:bgw.shape[0], :bgw.shape[1]] = bgw
        fr = bg.copy()
        o = 0
        for l in lines:
            k = int(l["x1"]) - int(l["x0"]) + 1
            fr[l["y"], l["x0"]:l["x1"] + 1] = px[o:o + k]; o += k
        s = trex_b200.DetectSettings(detect_threshold=8, detect_threshold_is_absolute=False, detect_size_filter=[(50, 100000)])
        bs = trex_b200.BackgroundSubtraction(bg, settings=s, max_batch=1)
        got = bs.apply([fr])
        assert len(got[0]) == 1 and got[0][0].num_pixels == len(px)
        segs, tail, head, _ = bs.midlines(0.5)[0]
        assert len(segs) > 2
        ratios.append(float(np.linalg.norm(np.diff(segs[:, :2], axis=0), axis=1).sum()) / float(g[f"b{i}_csv"][2]))
    ratios = np.array(ratios)
    assert abs(ratios.mean() - 1) < 0.03 and ratios.std() < 0.05
