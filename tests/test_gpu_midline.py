"""GPU parity of the posture chain (N4): tb_seg_midlines / tb_seg_posture vs the oracle's restatement of Outline::smooth,
offset_to_middle and calculate_midline (T/tracking/Outline.cpp:330-452,454-718,768-868; C/misc/CircularGraph.cpp:12-606),
Midline::post_process / normalize (:870-1456) and the posture / legacy crops (T/tracking/FilterCache.cpp:21-115,266-276).
Raw midlines are bit-exact: both sides evaluate the reference's float code without contraction.  Normalised midlines go through
atan2 / cos / sin in double on both sides (glibc vs CUDA libm, both within 1-2 double ulp): equal to the last float bit except
where a double result sits on a float rounding boundary -- asserted to 1e-5 absolute, and the bit-exact share is reported."""
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
    """1080p, 100 elongated individuals per frame: every blob's midline equals the oracle's, bit for bit -- pointy and broad tails."""
    import trex_b200
    from trex_b200.synthetic import BlobWorld
    world = BlobWorld(n_blobs=100, seed=6)
    frames = world.frames(2)
    bs = trex_b200.BackgroundSubtraction(world.bg, settings=trex_b200.DetectSettings(), max_batch=2)
    n, ok = _check(bs, frames, 1.0, every=3, peak_mode=1)
    assert n > 150 and ok > 0.25 * n
    n, ok = _check(bs, frames, 1.0)
    assert n > 150 and ok > 0.9 * n
    # an elongated ellipse: tail and head sit at opposite ends of the outline
    segs, tail, head, pts = bs.midlines(1.0)[0]
    assert tail == 0 and abs(head - len(pts) / 2) < len(pts) * 0.2 and len(segs) > 10


@pytest.mark.parametrize("rd,kw", [(0.5, {}), (1.0, dict(outline_approximate=0)), (1.0, dict(outline_smooth_samples=0, midline_invert=1)),
                                   (2.0, dict(outline_smooth_samples=2, outline_smooth_step=2, midline_start_with_head=1)),
                                   (1.0, dict(peak_mode=1)), (0.5, dict(peak_mode=1, outline_approximate=0)),
                                   (1.0, dict(peak_mode=1, outline_approximate=5, midline_start_with_head=1)), (1.0, dict(outline_approximate=8))])
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
        bs.midlines(1.0, peak_mode=2)
    with pytest.raises(trex_b200.TrexB200Error):
        bs.midlines(1.0, outline_approximate=9)
    with pytest.raises(trex_b200.TrexB200Error):
        bs.posture_async(1.0, midline_resolution=1)
    assert len(bs.midlines(1.0)) == 1 and len(bs.midlines(1.0, peak_mode=1)) == 1


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


def _posture_reference(blob, rd, P, move=None, fix=-1.0):
    """The oracle's chain for one blob: raw midline -> post_process -> normalize; None where the reference has no midline."""
    from oracle import posture, seg
    ol = seg.outline_resample(seg.longest_outline(blob.lines), rd)
    try:
        segs, tail, head, pts = posture.calculate_midline(ol, P)
    except ValueError:
        return None
    try:
        pp, t2, h2, inv = posture.post_process(segs, P, move_dir=move, tail=tail, head=head)
    except IndexError:
        return dict(segs=segs, tail=tail, head=head, threw=True)
    return dict(segs=segs, tail=tail, head=head, pp=pp, t2=t2, h2=h2, inv=inv, norm=posture.normalize(pp, P, fix))


@pytest.mark.parametrize("kw", [{}, dict(midline_resolution=12, midline_stiff_percentage=0.3, midline_start_with_head=1),
                                dict(peak_mode=1, midline_invert=1, midline_stiff_percentage=0.0)])
def test_posture_chain_async_normalised_midlines(kw):
    """tb_seg_posture enqueued BEFORE tb_seg_wait (blob count read on the device), fetch = 2: raw midlines bit-exact, normalised
    midlines (Individual::calculate_midline_for = post_process + normalize) against the oracle for every blob of the batch."""
    import torch
    import trex_b200
    from oracle import posture
    from trex_b200.synthetic import BlobWorld
    world = BlobWorld(n_blobs=100, seed=11)
    frames = world.frames(3)
    bs = trex_b200.BackgroundSubtraction(world.bg, settings=trex_b200.DetectSettings(), max_batch=3)
    bs.submit(np.ascontiguousarray(frames), fetch=1)
    bs.posture_async(1.0, fetch=2, **kw)               # no wait() in between
    bs.posture_wait()
    got = [bs.result(i) for i in range(3)]
    res = bs.posture_result()
    P = posture.default_params(**kw)
    RES = int(P.midline_resolution)
    n = res["n_blobs"]
    assert n == sum(len(g) for g in got) > 200 and res["midline_resolution"] == RES
    k = n_norm = exact = 0
    worst = 0.0
    for blobs in got:
        for b in blobs:
            ref = _posture_reference(b, 1.0, P)
            so, ns, tail, head = (int(v) for v in res["midlines"][k])
            nr = res["normalized"][k]
            if ref is None:
                assert ns == 0 and nr["n_points"] == 0
            else:
                assert (tail, head) == (ref["tail"], ref["head"]) and np.array_equal(res["segments"][so:so + ns], ref["segs"]), k
                if ref.get("threw"):
                    assert nr["n_points"] == 0 and nr["flags"] & 2
                elif ref["norm"] is None:
                    assert nr["n_points"] == 0 and nr["flags"] & 4, k
                else:
                    pts, length, angle, off = ref["norm"]
                    assert nr["n_points"] == RES and (nr["tail"], nr["head"]) == (ref["t2"], ref["h2"]) and (nr["flags"] & 1) == int(ref["inv"])
                    assert np.array_equal([nr["offx"], nr["offy"]], off) and nr["len"] == np.float32(length), k
                    g = res["norm_points"][k]
                    assert abs(float(nr["angle"]) - angle) < 1e-6, k
                    worst = max(worst, float(np.abs(g - pts).max()))
                    assert np.abs(g - pts).max() < 1e-4, k
                    exact += int(np.array_equal(g, pts) and nr["angle"] == np.float32(angle))
                    n_norm += 1
            k += 1
    print(f"normalised midlines: {n_norm} of {n} blobs, {exact} bit-exact, worst |d| {worst:.3g}")
    assert n_norm > 0.8 * n and exact > 0.9 * n_norm


def test_posture_movement_direction_and_fixed_length():
    """Per-blob device inputs: MovementInformation::direction (inverts midlines that point against the movement) and
    Midline::normalize(fix_length) as Individual::fixed_midline calls it."""
    import torch
    import trex_b200
    from oracle import posture
    from trex_b200.synthetic import BlobWorld
    world = BlobWorld(h=540, w=960, n_blobs=30, seed=3)
    frames = world.frames(2)
    bs = trex_b200.BackgroundSubtraction(world.bg, settings=trex_b200.DetectSettings(), max_batch=2)
    got = bs.apply(frames)
    n = sum(len(g) for g in got)
    rng = np.random.default_rng(0)
    move = rng.standard_normal((n, 2)).astype(np.float32); move[::5] = 0
    fix = (rng.random(n).astype(np.float32) * 60 + 20); fix[::3] = -1
    dev = torch.device("cuda", 0)
    move_d, fix_d = torch.from_numpy(move).to(dev), torch.from_numpy(fix).to(dev)
    bs.posture_async(1.0, fetch=1, move_direction_dev=move_d.data_ptr(), fix_length_dev=fix_d.data_ptr())
    bs.posture_wait()
    res = bs.posture_result()
    P = posture.default_params()
    k = checked = inverted = 0
    for blobs in got:
        for b in blobs:
            ref = _posture_reference(b, 1.0, P, move=move[k], fix=float(fix[k]))
            nr = res["normalized"][k]
            if ref is not None and not ref.get("threw") and ref["norm"] is not None:
                pts, length, angle, off = ref["norm"]
                assert nr["n_points"] == 25 and (nr["flags"] & 1) == int(ref["inv"]) and (nr["tail"], nr["head"]) == (ref["t2"], ref["h2"]), k
                assert abs(float(nr["len"]) - length) < 1e-3 and np.abs(res["norm_points"][k] - pts).max() < 1e-3, k
                checked += 1; inverted += int(ref["inv"])
            elif ref is not None and not ref.get("threw"):
                assert nr["n_points"] == 0
            k += 1
    assert checked > 0.7 * n and 0 < inverted < checked


@pytest.mark.parametrize("mode", ["posture", "legacy"])
def test_posture_normalised_crops(mode):
    """individual_image_normalization = posture (the reference's default) / legacy: the crops of the batch re-rendered through
    Midline::transform + normalize_image, byte-equal to the oracle's cv::warpAffine restatement fed with the device's own
    (angle, offset) -- which equal the oracle's to the bit for almost every blob, see above --, and chained into the CNN."""
    import trex_b200
    from oracle import posture, seg as oseg
    from trex_b200.synthetic import BlobWorld
    world = BlobWorld(h=540, w=960, n_blobs=40, seed=8)
    frames = world.frames(2)
    s = trex_b200.DetectSettings(individual_image_normalization=mode)
    bs = trex_b200.BackgroundSubtraction(world.bg, settings=s, max_batch=2, max_individuals=64)
    bs.submit(np.ascontiguousarray(frames), fetch=2)
    bs.posture_async(1.0, fetch=1, median_midline_length_px=42.0)
    bs.posture_wait()
    got = [bs.result(i) for i in range(2)]
    res = bs.posture_result()
    crops, idx = bs.crops()
    n = res["n_blobs"]
    assert len(crops) == n == len(res["crop_valid"]) and n > 50
    flat = [b for blobs in got for b in blobs]
    n_valid = 0
    for c in range(n):
        b, nr = flat[int(idx[c])], res["normalized"][int(idx[c])]
        if nr["n_points"] == 0:
            assert res["crop_valid"][c] == 0 and not crops[c].any()
            continue
        assert res["crop_valid"][c] == 1
        exp = posture.crop_blob_posture(b.lines, b.pixels, world.bg, oseg.DIFF_ABSOLUTE, float(nr["angle"]), (float(nr["offx"]), float(nr["offy"])),
                                        42.0, legacy=(mode == "legacy"))
        assert np.array_equal(crops[c], exp), c
        n_valid += 1
    assert n_valid > 0.8 * n
    # the network consumes the normalised crops from the device like any other crops
    net = trex_b200.VINetwork(16, max_images=2 * 64, precision="bf16x3")
    from trex_b200.weights import random_v118_3_state_dict
    net.load_weights(random_v118_3_state_dict(16, seed=0))
    probs = net.probabilities(crops[..., None])
    assert probs.shape == (n, 16) and np.allclose(probs.sum(1), 1, atol=1e-4)


def test_posture_of_thresholded_blobs_with_retries():
    """posture::calculate_posture (Posture.cpp:305-400) through tb_seg_posture_thresholded: graded blobs whose thresholded sub-blobs
    shrink round by round (several sub-blobs per parent, ties, parents that end without a midline) against the oracle's loop, blob by
    blob: the outline the midline saw, the segments, tail / head; the number of rounds equals the longest retry chain."""
    import trex_b200
    from oracle import posture, seg as oseg
    rng = np.random.default_rng(4)
    H, W = 160, 256
    bg = np.full((H, W), 200, np.uint8)
    fr = bg.copy()
    yy, xx = np.mgrid[0:H, 0:W]
    for k in range(14):                                   # elongated blobs with a darkness gradient, two of them with two dark cores
        cx, cy, a, b, th = rng.uniform(30, W - 30), rng.uniform(20, H - 20), rng.uniform(6, 22), rng.uniform(2, 6), rng.uniform(0, np.pi)
        u = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th); v = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
        r2 = (u / a) ** 2 + (v / b) ** 2
        m = r2 < 1
        depth = rng.uniform(20, 120)
        core = np.where(k % 5 == 0, np.minimum(((u - a / 2) / (a / 3)) ** 2 + (v / b) ** 2, ((u + a / 2) / (a / 3)) ** 2 + (v / b) ** 2), r2)
        fr[m] = np.clip(200 - depth * (1 - 0.9 * np.sqrt(core[m]).clip(0, 1)) - rng.integers(0, 4, int(m.sum())), 0, 199).astype(np.uint8)
    fr[5, 5:8] = 150; fr[150, 200] = 120                  # tiny blobs: no midline at any threshold
    frames = np.stack([fr, np.roll(fr, 7, axis=1)])
    det = trex_b200.BackgroundSubtraction(bg, settings=trex_b200.DetectSettings(detect_threshold=10, detect_size_filter=[]), max_batch=2)
    pst = trex_b200.BackgroundSubtraction(bg, settings=trex_b200.DetectSettings(detect_threshold=0, detect_size_filter=[]), max_batch=2)
    got = det.apply(frames)
    T0 = 12
    rounds, res = det.posture_thresholded(pst, track_posture_threshold=T0, outline_resample=1.0, fetch=2)
    flat = [b for blobs in got for b in blobs]
    assert res["n_blobs"] == len(flat) > 20
    exp_rounds, n_mid, n_outline_only, n_none = 1, 0, 0, 0
    for k, b in enumerate(flat):
        so, ns, tail, head = (int(v) for v in res["midlines"][k])
        ro, rn = int(res["outlines"][k][2]), int(res["outlines"][k][3])
        try:
            ref = posture.calculate_posture(b.lines, b.pixels, bg, track_posture_threshold=T0, outline_resample=1.0, method=oseg.DIFF_ABSOLUTE)
        except ValueError:
            assert ns == 0 and rn == 0, k
            n_none += 1
            continue
        if ref["segments"] is None:
            assert ns == 0 and np.array_equal(res["points"][ro:ro + rn], ref["outline"]), k
            n_outline_only += 1
        else:
            assert (tail, head) == (ref["tail"], ref["head"]), k
            assert np.array_equal(res["points"][ro:ro + rn], ref["outline"]) and np.array_equal(res["segments"][so:so + ns], ref["segments"]), k
            exp_rounds = max(exp_rounds, (ref["threshold"] - T0) // 2 + 1)
            nm = posture.normalize(posture.post_process(ref["segments"], tail=ref["tail"], head=ref["head"])[0])
            nr = res["normalized"][k]
            assert (nm is None and nr["n_points"] == 0) or (nm is not None and np.abs(res["norm_points"][k] - nm[0]).max() < 1e-4), k
            n_mid += 1
    print(f"rounds {rounds}, midlines {n_mid}, outline only {n_outline_only}, nothing {n_none}")
    assert n_mid > 15 and n_outline_only + n_none >= 2 and rounds >= exp_rounds and rounds > 1
