"""GPU parity tests of the segmentation path: CUDA (through the C ABI) vs the CPU oracle and the
committed golden vectors of the reference (videos/test.pv).  Bit-exact: integer work."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

PV = dict(detect_threshold=9, detect_size_filter=[(1, 10000)], cm_per_pixel=1.0)


def _mk(bg, max_batch=4, max_individuals=0, dense=False, **kw):
    import trex_b200
    s = trex_b200.DetectSettings(**kw)
    h, w = bg.shape
    cap = dict(max_runs_per_frame=h * w // 2 + 16, max_pixels_per_frame=h * w) if dense else {}
    return trex_b200.BackgroundSubtraction(bg, settings=s, max_batch=max_batch, max_individuals=max_individuals, **cap)


def _as_list(blobs):
    return [(b.lines.tobytes(), b.pixels.tobytes()) for b in blobs]


def _oracle(frame, bg, **kw):
    from oracle import seg
    keys = {k: v for k, v in kw.items() if k in seg.Params.__dataclass_fields__}
    return seg.segment_frame(frame, bg, seg.Params(**keys))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "testpv_golden.npz"))


def test_reference_fixture_full_frames(gold):
    """The reference's own output for videos/test_frames 0 and 100 (2304x2304), as a set and via bid."""
    from oracle import seg
    bs = _mk(gold["average"], max_batch=2, **PV)
    frames = [gold["full0_frame"], gold["full100_frame"]]
    got = bs.apply(frames)
    for idx, g in zip((0, 100), got):
        ref = seg.Blobs(gold[f"full{idx}_lines"], gold[f"full{idx}_pixels"], gold[f"full{idx}_line_off"], gold[f"full{idx}_px_off"])
        assert len(g) == len(ref)
        assert set(_as_list(g)) == ref.as_set()
        canon = seg.segment_frame(gold[f"full{idx}_frame"], gold["average"], seg.Params(**PV))
        assert _as_list(g) == canon.as_list()            # canonical order too
        for b in g:
            assert b.bid == seg.blob_id(b.lines)
            assert b.bounds == (int(b.lines["x0"].min()), int(b.lines["y"].min()), int(b.lines["x1"].max()), int(b.lines["y"].max()))


def test_reference_fixture_windows(gold):
    from oracle import seg
    for (i, y0, x0) in gold["windows"]:
        fr = gold[f"win{i}_frame"]
        h, w = fr.shape
        bs = _mk(gold["average"][y0:y0 + h, x0:x0 + w], max_batch=1, **PV)
        (g,) = bs.apply([fr])
        ref = seg.Blobs(gold[f"win{i}_lines"], gold[f"win{i}_pixels"], gold[f"win{i}_line_off"], gold[f"win{i}_px_off"])
        inner = {(b.lines.tobytes(), b.pixels.tobytes()) for b in g
                 if b.bounds[0] > 0 and b.bounds[1] > 0 and b.bounds[2] < w - 1 and b.bounds[3] < h - 1}
        assert inner == ref.as_set()
        bs.deinit()


def test_synthetic_1080p_vs_oracle():
    """BASELINE config 2: synthetic 1920x1080, 100 moving blobs, bit-exact masks/lines/pixels."""
    from trex_b200.synthetic import BlobWorld
    world = BlobWorld(n_blobs=100, seed=1234)
    frames = world.frames(6)
    kw = dict(detect_threshold=15, detect_size_filter=[(10, 100000)])
    bs = _mk(world.bg, max_batch=4, **kw)
    got = bs.apply(frames)       # two submits (4 + 2)
    assert len(got) == 6
    for f in range(6):
        ref = _oracle(frames[f], world.bg, **kw)
        assert _as_list(got[f]) == ref.as_list(), f
        assert len(ref) >= 90
    from oracle import seg
    assert np.array_equal(bs.debug_binary(frames[0]), seg.generate_binary(frames[0], world.bg, seg.Params(**kw)))


@pytest.mark.parametrize("kw", [
    dict(detect_threshold=15),
    dict(detect_threshold=15, detect_threshold_is_absolute=False),
    dict(detect_threshold=-15),
    dict(detect_threshold=10, threshold_maximum=60),
    dict(detect_threshold=40, enable_difference=False),
    dict(detect_threshold=15, image_invert=True),
    dict(detect_threshold=0),
    dict(detect_threshold=255),
])
def test_threshold_variants(kw):
    """The generate_binary branch family (RawProcessing.cpp:364-399,529-539) on a noisy frame."""
    rng = np.random.default_rng(11)
    bg = rng.integers(90, 160, (120, 208)).astype(np.uint8)
    fr = np.clip(bg.astype(int) + rng.integers(-40, 41, bg.shape), 0, 255).astype(np.uint8)
    fr[30:60, 50:120] = 10
    fr[40:45, 70:80] = 0
    kw = dict(kw, detect_size_filter=[])
    bs = _mk(bg, max_batch=1, dense=True, **kw)
    (g,) = bs.apply([fr])
    ref = _oracle(fr, bg, **kw)
    assert _as_list(g) == ref.as_list()
    from oracle import seg
    keys = {k: v for k, v in kw.items() if k in seg.Params.__dataclass_fields__}
    assert np.array_equal(bs.debug_binary(fr), seg.generate_binary(fr, bg, seg.Params(**keys)))


def test_adversarial_geometry():
    """SURVEY s8d config 2 adversarial cases + generic widths (not a multiple of 16)."""
    for (h, w) in ((64, 96), (37, 53), (200, 1000), (5, 16), (1, 1), (300, 17)):
        rng = np.random.default_rng(h * 1000 + w)
        bg = np.full((h, w), 100, np.uint8)
        fr = bg.copy()
        m = rng.random((h, w)) < 0.35                      # dense noise: many merges, diagonal contacts
        fr[m] = 200
        fr[rng.random((h, w)) < 0.02] = 0                  # grey 0 is never foreground
        fr[h - 1, w - 1] = 200; fr[0, 0] = 200; fr[h - 1, 0] = 200; fr[0, w - 1] = 200
        kw = dict(detect_threshold=15, detect_size_filter=[])
        bs = _mk(bg, max_batch=1, dense=True, **kw)
        (g,) = bs.apply([fr])
        ref = _oracle(fr, bg, **kw)
        assert _as_list(g) == ref.as_list(), (h, w)
        bs.deinit()
    # empty frame, full frame, spiral (one blob with many merges), checkerboard (8-conn: one blob)
    h, w = 96, 160
    bg = np.full((h, w), 100, np.uint8)
    cases = {"empty": bg.copy(), "full": np.full((h, w), 200, np.uint8)}
    cb = bg.copy(); yy, xx = np.mgrid[0:h, 0:w]; cb[(yy + xx) % 2 == 0] = 200; cases["checker"] = cb
    sp = bg.copy()
    for k in range(0, 40, 4):
        sp[k, k:w - k] = 200; sp[h - 1 - k, k:w - k] = 200; sp[k:h - k, w - 1 - k] = 200; sp[k + 4:h - k, k] = 200
    cases["spiral"] = sp
    comb = bg.copy(); comb[10, :] = 200; comb[10:80, ::2] = 200; cases["comb"] = comb
    bs = _mk(bg, max_batch=8, dense=True, detect_threshold=15, detect_size_filter=[])
    got = bs.apply(list(cases.values()))
    for (name, fr), g in zip(cases.items(), got):
        ref = _oracle(fr, bg, detect_threshold=15, detect_size_filter=[])
        assert _as_list(g) == ref.as_list(), name
    assert len(got[0]) == 0 and len(got[1]) == 1 and len(got[2]) == 1


def test_size_filter_and_cm_per_pixel():
    bg = np.full((64, 256), 100, np.uint8)
    fr = bg.copy()
    for i, n in enumerate((9, 10, 11, 40, 41)):
        fr[4 + 8 * i, 8:8 + n] = 30
    for kw in (dict(detect_size_filter=[(10, 41)]), dict(detect_size_filter=[(2.5, 10.0)], cm_per_pixel=0.5),
               dict(detect_size_filter=[(9, 10), (40, 41)]), dict(detect_size_filter=[])):
        bs = _mk(bg, max_batch=1, detect_threshold=15, **kw)
        (g,) = bs.apply([fr])
        ref = _oracle(fr, bg, detect_threshold=15, **kw)
        assert _as_list(g) == ref.as_list(), kw
        bs.deinit()


def test_crops_vs_oracle():
    from oracle import seg
    from trex_b200.synthetic import BlobWorld
    world = BlobWorld(h=540, w=960, n_blobs=40, seed=3, semi=(22, 6))
    frames = world.frames(3)
    big = world.bg.copy(); big[100:300, 200:420] = 20          # a blob larger than the crop: centre crop
    frames = np.concatenate([frames, big[None]])
    for method, kw in ((seg.DIFF_ABSOLUTE, {}), (seg.DIFF_NONE, dict(track_background_subtraction=False)),
                       (seg.DIFF_SIGN, dict(track_threshold_is_absolute=False))):
        bs = _mk(world.bg, max_batch=4, max_individuals=64, detect_threshold=15, detect_size_filter=[(10, 100000)], **kw)
        got = bs.apply(frames)
        crops, idx = bs.crops()
        n = 0
        for f in range(len(frames)):
            ref = _oracle(frames[f], world.bg, detect_threshold=15, detect_size_filter=[(10, 100000)])
            assert _as_list(got[f]) == ref.as_list()
            for k in range(min(len(ref), 64)):
                assert np.array_equal(crops[n], seg.crop_blob(*ref.blob(k), world.bg, method)), (method, f, k)
                n += 1
        assert n == len(crops)
        bs.deinit()


@pytest.mark.parametrize("scale,method", [(0.5, "absolute"), (1.5, "absolute"), (2.5, "sign"), (0.75, "none"), (1.1, "absolute")])
def test_scaled_crops_vs_oracle(scale, method):
    """individual_image_scale != 1 (FilterCache.cpp:178-180): nearest-neighbour resize of the masked blob image, then the centre
    pad / centre crop; scale 2.5 makes the benchmark's blobs larger than 80x80 (crop branch)."""
    from oracle import seg
    from trex_b200.synthetic import BlobWorld
    world = BlobWorld(h=272, w=480, n_blobs=14, seed=9, margin=30)
    frames = world.frames(3)
    kw = dict(detect_threshold=15, detect_size_filter=[(1, 100000)], individual_image_scale=scale,
              track_background_subtraction=method != "none", track_threshold_is_absolute=method != "sign")
    bs = _mk(world.bg, max_batch=4, max_individuals=32, **kw)
    got = bs.apply(frames)
    crops, idx = bs.crops()
    m = {"none": seg.DIFF_NONE, "absolute": seg.DIFF_ABSOLUTE, "sign": seg.DIFF_SIGN}[method]
    n = 0
    for f in range(3):
        ref = _oracle(frames[f], world.bg, **kw)
        assert _as_list(got[f]) == ref.as_list()
        for k in range(min(len(ref), 32)):
            assert np.array_equal(crops[n], seg.crop_blob_scaled(*ref.blob(k), world.bg, m, scale)), (f, k)
            n += 1
    assert n == len(crops) and n > 20 and crops.any()


def test_idempotence_render_relabel():
    """test_matching.cpp:1556-1602 property on the GPU path: render blobs -> relabel -> same lines."""
    from trex_b200.synthetic import BlobWorld
    world = BlobWorld(h=272, w=480, n_blobs=15, seed=9, margin=30)
    fr = world.frame()
    kw = dict(detect_threshold=15, detect_size_filter=[])
    bs = _mk(world.bg, max_batch=1, **kw)
    (g,) = bs.apply([fr])
    render = np.zeros_like(fr)
    for b in g:
        o = 0
        for l in b.lines:
            n = int(l["x1"]) - int(l["x0"]) + 1
            render[l["y"], l["x0"]:l["x1"] + 1] = b.pixels[o:o + n]; o += n
    bs2 = _mk(np.zeros_like(fr), max_batch=1, detect_threshold=0, enable_difference=False, detect_size_filter=[])
    (g2,) = bs2.apply([render])
    assert _as_list(g2) == _as_list(g)


def test_capacity_and_state_errors():
    import trex_b200
    bg = np.full((64, 64), 100, np.uint8)
    bs = trex_b200.BackgroundSubtraction(width=64, height=64, max_batch=1)
    with pytest.raises(trex_b200.TrexB200Error) as e:
        bs.apply([bg])                       # no background yet: the reference's pipeline stays paused
    assert e.value.code == -3
    bs.set_background(bg)
    with pytest.raises(trex_b200.TrexB200Error):
        bs.update_settings(trex_b200.DetectSettings(use_closing=True, closing_size=9))     # element larger than 15x15
    # a rejected tb_seg_set_params leaves the previous settings fully in place: the handle still segments
    for bad in (dict(use_adaptive_threshold=True, adaptive_threshold_scale=99.0), dict(blur_difference=True, use_closing=True, closing_size=9)):
        with pytest.raises(trex_b200.TrexB200Error):
            bs.update_settings(trex_b200.DetectSettings(**bad))
    fr0 = bg.copy(); fr0[10:20, 10:30] = 30
    assert [b.num_pixels for b in bs.apply([fr0])[0]] == [200]
    bs2 = trex_b200.BackgroundSubtraction(bg, max_batch=1, max_runs_per_frame=16,
                                          settings=trex_b200.DetectSettings(detect_size_filter=[]))
    fr = bg.copy(); fr[::2, ::2] = 200       # 1024 runs > capacity
    with pytest.raises(trex_b200.TrexB200Error) as e:
        bs2.apply([fr])
    assert e.value.code == -4


def test_config4_256_individuals_and_config5_4k():
    """BASELINE configs[3] / [4] geometry: 256 blobs at 1080p and 100 blobs at 3840x2160, bit-exact vs oracle."""
    from trex_b200.synthetic import BlobWorld
    kw = dict(detect_threshold=15, detect_size_filter=[(10, 100000)])
    for (h, w, n, frames) in ((1080, 1920, 256, 2), (2160, 3840, 100, 2)):
        world = BlobWorld(h=h, w=w, n_blobs=n, seed=77)
        fr = world.frames(frames)
        bs = _mk(world.bg, max_batch=frames, max_individuals=256, **kw)
        got = bs.apply(fr)
        crops, idx = bs.crops()
        from oracle import seg
        k = 0
        for f in range(frames):
            ref = _oracle(fr[f], world.bg, **kw)
            assert _as_list(got[f]) == ref.as_list(), (h, w, f)
            assert len(ref) >= int(0.85 * n)
            for j in range(min(len(ref), 256)):
                if j % 37 == 0:
                    assert np.array_equal(crops[k + j], seg.crop_blob(*ref.blob(j), world.bg, seg.DIFF_ABSOLUTE))
            k += min(len(ref), 256)
        assert k == len(crops)
        bs.deinit()


@pytest.mark.parametrize("kw", [
    dict(use_closing=True, closing_size=3),
    dict(use_closing=True, closing_size=2),
    dict(dilation_size=3),
    dict(dilation_size=2),
    dict(use_closing=True, closing_size=1, dilation_size=2),
    dict(dilation_size=-3),
    dict(use_closing=True, closing_size=2, dilation_size=-2),
    dict(open_size=2),                                           # north_star's "2x2 morphological open" (not a reference stage; default off)
    dict(open_size=3, use_closing=True, closing_size=2),
    dict(open_size=2, dilation_size=2),
])
def test_morphology_vs_oracle(kw):
    """Optional closing / dilation of generate_binary (RawProcessing.cpp:438-550); the oracle's closing and
    positive dilation are themselves checked against OpenCV in tests/test_oracle_golden.py."""
    from oracle import seg
    rng = np.random.default_rng(21)
    bg = np.full((144, 208), 120, np.uint8)
    fr = bg.copy()
    fr[rng.random(fr.shape) < 0.06] = 30
    fr[40:70, 50:120] = 20; fr[50:55, 70:90] = 120; fr[100:103, 10:200:3] = 200
    fr[rng.random(fr.shape) < 0.01] = 0
    kw = dict(kw, detect_threshold=15, detect_size_filter=[])
    bs = _mk(bg, max_batch=2, dense=True, **kw)
    got = bs.apply([fr, bg])
    ref = _oracle(fr, bg, **kw)
    assert _as_list(got[0]) == ref.as_list()
    assert _as_list(got[1]) == _oracle(bg, bg, **kw).as_list()
    keys = {k: v for k, v in kw.items() if k in seg.Params.__dataclass_fields__}
    assert np.array_equal(bs.debug_binary(fr), seg.generate_binary(fr, bg, seg.Params(**keys)))


@pytest.mark.parametrize("kw,size", [
    (dict(blur_difference=True), (144, 208)),
    (dict(blur_difference=True, detect_threshold_is_absolute=False, image_invert=True), (97, 131)),
    (dict(blur_difference=True), (540, 960)),
    (dict(use_adaptive_threshold=True, adaptive_threshold_scale=0.1), (144, 208)),
    (dict(use_adaptive_threshold=True), (144, 208)),                                     # default scale 2: neighbourhood 2 * cols + 1
    (dict(use_adaptive_threshold=True, adaptive_threshold_scale=0.02, detect_threshold=-4), (97, 131)),
    (dict(use_adaptive_threshold=True, adaptive_threshold_scale=0.05, use_closing=True, closing_size=2), (144, 208)),
    (dict(use_adaptive_threshold=True, adaptive_threshold_scale=0.07, dilation_size=-3), (144, 208)),
    (dict(use_adaptive_threshold=True, adaptive_threshold_scale=0.03, enable_difference=False, detect_threshold=30), (540, 960)),
])
def test_blur_difference_and_adaptive_threshold_vs_oracle(kw, size):
    """generate_binary's blur_difference (RawProcessing.cpp:371-387) and use_adaptive_threshold (:427-434,487,526) stages; the
    oracle's versions are checked against the OpenCV calls in tests/test_oracle_golden.py.  11 frames: the box filter works in
    sub-batches of 8."""
    from oracle import seg
    h, w = size
    rng = np.random.default_rng(h)
    bg = rng.integers(110, 130, (h, w)).astype(np.uint8)
    frames = []
    for f in range(11):
        fr = np.clip(bg.astype(int) + rng.integers(-7, 8, bg.shape), 1, 255).astype(np.uint8)
        y, x = int(rng.integers(0, h - 40)), int(rng.integers(0, w - 60))
        fr[y:y + 30, x:x + 50] = 20; fr[y + 10:y + 14, x + 10:x + 30] = 120
        fr[0:9, w - 17:w] = 30; fr[h - 3:h, 0:40] = 250
        fr[rng.random(fr.shape) < 0.003] = 0
        frames.append(fr)
    kw = dict(dict(detect_threshold=9, detect_size_filter=[]), **kw)
    bs = _mk(bg, max_batch=11, dense=True, **kw)
    got = bs.apply(frames)
    keys = {k: v for k, v in kw.items() if k in seg.Params.__dataclass_fields__}
    n_blobs = 0
    for f in (0, 5, 8, 10):
        ref = _oracle(frames[f], bg, **kw)
        assert _as_list(got[f]) == ref.as_list(), f
        n_blobs += len(ref)
    assert n_blobs > 0
    assert np.array_equal(bs.debug_binary(frames[3]), seg.generate_binary(frames[3], bg, seg.Params(**keys)))


def test_blur_difference_errors():
    import trex_b200
    with pytest.raises(trex_b200.TrexB200Error):      # the 25x25 window does not fit
        _mk(np.zeros((12, 64), np.uint8), blur_difference=True)
    with pytest.raises(trex_b200.TrexB200Error):
        _mk(np.zeros((64, 64), np.uint8), use_adaptive_threshold=True, adaptive_threshold_scale=-1.0)


def test_pv_file_from_gpu_results(tmp_path, gold):
    """GPU blobs -> PV15 file (trex_b200.pv_writer) -> oracle reader: same blobs as the reference's file."""
    import trex_b200
    from oracle import seg
    from trex_b200.pv_writer import PVWriter
    bs = _mk(gold["average"], max_batch=2, **PV)
    bs.apply([gold["full0_frame"], gold["full100_frame"]], materialize=False)
    path = str(tmp_path / "gpu.pv")
    with PVWriter(path, 2304, 2304, gold["average"], name="gpu") as w:
        for i, idx in enumerate((0, 100)):
            w.add_result(bs, i, timestamp_us=idx * 40000, source_index=idx)
    try:
        from oracle.pv15 import PV15
        pv = PV15(path)
    except FileNotFoundError:
        pytest.skip("oracle/_ref/libminilzo.so not available")
    for i, idx in enumerate((0, 100)):
        ref = seg.Blobs(gold[f"full{idx}_lines"], gold[f"full{idx}_pixels"], gold[f"full{idx}_line_off"], gold[f"full{idx}_px_off"])
        assert pv.frame(i).as_set() == ref.as_set()


@pytest.mark.parametrize("trk_kw,method", [
    (dict(detect_threshold=40), 1),                                                  # absolute, >=
    (dict(detect_threshold=40, detect_threshold_is_absolute=False), 2),              # sign
    (dict(detect_threshold=60, enable_difference=False), 0),                         # none: grey value >= T
    (dict(detect_threshold=15), 1),                                                  # T2 == T1: >= vs > differs only outside blobs
])
def test_tracker_side_rethreshold(trk_kw, method):
    """pixel::threshold_blob on every detection blob (N3a): GPU vs oracle, bit-exact, incl. crops of tracker-side blobs."""
    from oracle import seg
    from trex_b200.synthetic import BlobWorld
    world = BlobWorld(h=540, w=960, n_blobs=40, seed=12, semi=(26, 9))
    frames = world.frames(3).copy()
    rng = np.random.default_rng(5)
    for f in frames:                                   # texture inside the blobs so that a higher threshold splits them
        m = np.abs(f.astype(int) - world.bg.astype(int)) > 15
        f[m] = np.clip(f[m].astype(int) + rng.integers(0, 70, int(m.sum())), 1, 255).astype(np.uint8)
    det_kw = dict(detect_threshold=15, detect_size_filter=[(10, 100000)])
    det = _mk(world.bg, max_batch=4, **det_kw)
    trk = _mk(world.bg, max_batch=4, max_individuals=64, detect_size_filter=[], **trk_kw)
    det.apply(frames, materialize=False)
    got = det.rethreshold(trk, fetch=2)
    crops, _ = trk.crops()
    c = 0
    for f in range(3):
        parents = _oracle(frames[f], world.bg, **det_kw)
        ref = seg.rethreshold(parents, world.bg, trk_kw["detect_threshold"], method)
        exp = sorted(ref.as_list(), key=lambda lp: tuple(np.frombuffer(lp[0], seg.LINE_DTYPE)[["y", "x0"]][0]))
        assert _as_list(got[f]) == exp, f
        assert len(ref) >= len(parents) or method != 1
        for k in range(min(len(got[f]), 64)):
            b = got[f][k]
            want = seg.crop_blob(b.lines, b.pixels, world.bg, seg.DIFF_ABSOLUTE)
            assert np.array_equal(crops[c + k], want)
        c += min(len(got[f]), 64)


@pytest.mark.parametrize("case", ["many_small_blobs", "runs_at_smem_limit", "runs_beyond_smem_limit", "tall_frame"])
def test_labeling_kernel_paths(case):
    """K2's shared-memory fast path (<= 8192 runs, <= 4352 rows; per-blob statistics in shared memory up to 512
    components, global atomics beyond) and the general path must agree with the oracle on either side of the limits."""
    rng = np.random.default_rng(42)
    if case == "tall_frame":
        h, w = 4400, 64                                   # more rows than the fast path's row table
    else:
        h, w = 512, 768
    bg = np.full((h, w), 100, np.uint8)
    fr = bg.copy()
    if case == "many_small_blobs":                        # ~1500 components, ~1500 runs: statistics via global atomics
        ys, xs = np.mgrid[2:h - 2:8, 2:w - 2:32]
        fr[ys.ravel(), xs.ravel()] = 20
        fr[4::16, 5:40] = 30                              # a few longer runs
    elif case == "runs_at_smem_limit":                    # exactly 8192 runs
        pts = [(y, x) for y in range(0, h, 2) for x in range(0, w, 3)][:8192]
        for y, x in pts:
            fr[y, x] = 20
    elif case == "runs_beyond_smem_limit":                # 8193 + noise: general path
        pts = [(y, x) for y in range(0, h, 2) for x in range(0, w, 3)][:8193]
        for y, x in pts:
            fr[y, x] = 20
        fr[1::2, ::5] = 10                                # joins runs vertically / diagonally
    else:
        for _ in range(300):
            y, x = int(rng.integers(0, h - 6)), int(rng.integers(0, w - 6))
            fr[y:y + int(rng.integers(1, 6)), x:x + int(rng.integers(1, 6))] = int(rng.integers(1, 60))
    kw = dict(detect_threshold=15, detect_size_filter=[(1, 1000000)])
    bs = _mk(bg, max_batch=2, dense=True, **kw)
    got = bs.apply([fr, bg])
    ref = _oracle(fr, bg, **kw)
    if case == "runs_at_smem_limit":
        assert bs.frame_info(0).n_runs == 8192
    assert len(ref) > 0 and _as_list(got[0]) == ref.as_list()
    assert got[1] == []


@pytest.mark.parametrize("method", ["absolute", "sign", "none"])
def test_moments_normalised_crops(method):
    """individual_image_normalization = moments (FilterCache.cpp:329-341): orientation from the blob's second moments,
    cv::warpAffine into the 80x80 canvas.  Byte-equal to the oracle (itself equal to the cv2.warpAffine call)."""
    import trex_b200
    from oracle import seg
    from trex_b200.synthetic import BlobWorld
    world = BlobWorld(h=400, w=640, n_blobs=30, seed=31, margin=30)
    frames = world.frames(3)
    # adversarial blobs: a single pixel, a long thin line (wider than the canvas), a big square, a blob touching the border
    frames[0][5, 7] = 10
    frames[0][200, 100:330] = 20
    frames[1][50:170, 300:420] = 30
    frames[2][0:9, 0:30] = 25
    kw = dict(detect_threshold=15, detect_size_filter=[(1, 1000000)])
    tbs = dict(track_background_subtraction=method != "none", track_threshold_is_absolute=method != "sign")
    s = trex_b200.DetectSettings(individual_image_normalization="moments", **kw, **tbs)
    bs = trex_b200.BackgroundSubtraction(world.bg, settings=s, max_batch=4, max_individuals=64)
    got = bs.apply(frames)
    crops, idx = bs.crops()
    m = {"absolute": seg.DIFF_ABSOLUTE, "sign": seg.DIFF_SIGN, "none": seg.DIFF_NONE}[method]
    n = 0
    for f in range(len(frames)):
        ref = _oracle(frames[f], world.bg, **kw)
        assert _as_list(got[f]) == ref.as_list()
        for k in range(min(len(ref), 64)):
            exp = seg.crop_blob_moments(*ref.blob(k), world.bg, m)
            assert np.array_equal(crops[n], exp), (f, k, int(np.abs(crops[n].astype(int) - exp).max()))
            n += 1
    assert n == len(crops) and n > 60


@pytest.mark.parametrize("diff,absolute", [(True, True), (True, False), (False, True)])
def test_recount_against_oracle(diff, absolute):
    """tb_seg_recount = pv::Blob::recount(threshold, background) (PVBlob.cpp:934-1027, Background.h:430-489) for every blob of a batch:
    the three difference methods, several thresholds incl. 0 (= num_pixels), cm_per_pixel != 1."""
    import trex_b200
    from oracle import seg as oseg
    from trex_b200.synthetic import BlobWorld
    world = BlobWorld(h=272, w=480, n_blobs=20, seed=9, margin=30)
    frames = world.frames(2)
    s = trex_b200.DetectSettings(enable_difference=diff, detect_threshold_is_absolute=absolute, cm_per_pixel=0.25,
                                 detect_threshold=15 if diff else 40, detect_size_filter=[])
    bs = trex_b200.BackgroundSubtraction(world.bg, settings=s, max_batch=2)
    got = bs.apply(frames)
    flat = [b for blobs in got for b in blobs]
    assert len(flat) > (10 if diff else 1)              # without the difference the bright background is one frame-wide blob
    method = oseg.DIFF_NONE if not diff else (oseg.DIFF_ABSOLUTE if absolute else oseg.DIFF_SIGN)
    for T in (0, 20, 60, 200):
        rc = bs.recount(T)
        exp = np.array([oseg.blob_recount(b.lines, b.pixels, world.bg, T, method, cm_per_pixel=0.25) for b in flat], np.float32)
        assert np.array_equal(rc, exp), T
    assert np.array_equal(bs.recount(0), np.array([b.num_pixels for b in flat], np.float32) * np.float32(0.0625))


def test_moments_crop_of_a_blob_with_more_than_1000_runs():
    """pv::Blob::calculate_moments sums a blob of more than 1000 runs in four packages with their own float accumulators (PVBlob.cpp:118-204), which changes
    the rounding of the sums and with it the orientation; the oracle follows that (pinned on the compiled PVBlob.cpp, tests/test_oracle_ref_pvblob.py) and so
    does blob_moments_kernel.  A porous ellipse of ~1600 runs next to ordinary blobs: crops byte-equal to the oracle's."""
    import trex_b200
    from oracle import seg
    rng = np.random.default_rng(41)
    H, W = 720, 960
    bg = np.full((H, W), 180, np.uint8)
    fr = bg.copy()
    yy, xx = np.mgrid[0:H, 0:W]
    u = (xx - 480) * np.cos(0.5) + (yy - 360) * np.sin(0.5); v = -(xx - 480) * np.sin(0.5) + (yy - 360) * np.cos(0.5)
    m = ((u / 330) ** 2 + (v / 90) ** 2 < 1) & (rng.random((H, W)) < 0.985)
    fr[m] = rng.integers(20, 120, int(m.sum())).astype(np.uint8)
    fr[30:50, 40:90] = 60; fr[650:660, 800:930] = 90
    kw = dict(detect_threshold=15, detect_size_filter=[(1, 10000000)])
    s = trex_b200.DetectSettings(individual_image_normalization="moments", **kw)
    bs = trex_b200.BackgroundSubtraction(bg, settings=s, max_batch=1, max_individuals=64, max_runs_per_frame=H * W // 8, max_pixels_per_frame=H * W)
    (g,) = bs.apply(fr[None])
    crops, _ = bs.crops()
    ref = _oracle(fr, bg, **kw)
    assert _as_list(g) == ref.as_list()
    assert max(len(ref.blob(k)[0]) for k in range(len(ref))) > 1000
    n = 0
    for k in range(min(len(ref), 64)):
        assert np.array_equal(crops[n], seg.crop_blob_moments(*ref.blob(k), bg, seg.DIFF_ABSOLUTE)), k
        n += 1
    assert n == len(crops) and n >= 3
