"""GPU parity of the outline stage (N4, first stage): tb_seg_outlines vs the oracle's literal emulation of
pixel::find_outer_points (C/processing/PixelTree.cpp:497-1130) and Outline::resample (T/tracking/Outline.cpp:724-766).
Bit-exact: raw points are half-integers, the resampling arithmetic is compiled without contraction on both sides."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(bs, frames, rd, every=1):
    from oracle import seg
    got = bs.apply(frames)
    raw, res = bs.outlines(rd)
    n = sum(len(g) for g in got)
    assert len(raw) == len(res) == n and n > 0
    k = 0
    for blobs in got:
        for b in blobs:
            if k % every == 0:
                ref = seg.longest_outline(b.lines)
                assert np.array_equal(raw[k], ref), k
                assert np.array_equal(res[k], seg.outline_resample(ref, rd)), k
            k += 1
    return n


@pytest.mark.parametrize("rd", [1.0, 0.5, 2.5, 0.0])
def test_outlines_of_noisy_blobs(rd):
    """Dense random images: blobs with holes, corner contacts, single pixels; every blob checked."""
    import trex_b200
    rng = np.random.default_rng(int(rd * 10))
    h, w = 96, 160
    bg = np.zeros((h, w), np.uint8)
    frames = [np.where(rng.random((h, w)) < d, 200, 0).astype(np.uint8) for d in (0.25, 0.5, 0.62, 0.8)]
    s = trex_b200.DetectSettings(detect_threshold=15, detect_size_filter=[])
    bs = trex_b200.BackgroundSubtraction(bg, settings=s, max_batch=4, max_runs_per_frame=h * w // 2 + 16, max_pixels_per_frame=h * w)
    assert _check(bs, frames, rd) > 300


def test_outlines_of_the_benchmark_workload():
    """1080p, 100 individuals per frame (BASELINE config 2 / 5 geometry), default outline_resample = 1."""
    import trex_b200
    from trex_b200.synthetic import BlobWorld
    world = BlobWorld(n_blobs=100, seed=4)
    frames = world.frames(3)
    bs = trex_b200.BackgroundSubtraction(world.bg, settings=trex_b200.DetectSettings(), max_batch=4)
    n = _check(bs, frames, 1.0, every=3)
    assert n > 250
    raw, res = bs.outlines(0.5)                             # videos/test.settings: outline_resample = 0.5
    assert all(len(b) >= len(a) for a, b in zip(raw, res) if len(a) > 8)


def test_outlines_of_large_blobs_row_table_path():
    """Blobs whose bounding box does not fit a bit image (or the CTA's pool) are traced through the row table: a frame-wide
    dense blob with holes, next to small ones that share its CTA."""
    import trex_b200
    rng = np.random.default_rng(12)
    h, w = 400, 640
    bg = np.zeros((h, w), np.uint8)
    fr = np.zeros((h, w), np.uint8)
    fr[20:380, 16:620] = np.where(rng.random((360, 604)) < 0.93, 180, 0)        # one huge blob full of holes
    fr[2:10, 2:30] = 99; fr[390:398, 100:140:3] = 77
    fr2 = np.where(rng.random((h, w)) < 0.35, 150, 0).astype(np.uint8)
    s = trex_b200.DetectSettings(detect_threshold=15, detect_size_filter=[])
    bs = trex_b200.BackgroundSubtraction(bg, settings=s, max_batch=2, max_runs_per_frame=h * w // 2 + 16, max_pixels_per_frame=h * w)
    n = _check(bs, [fr, fr2], 1.0)
    assert n > 1000


def test_outlines_errors_and_empty_batch():
    import trex_b200
    bg = np.full((64, 64), 100, np.uint8)
    bs = trex_b200.BackgroundSubtraction(bg, max_batch=2)
    with pytest.raises(trex_b200.TrexB200Error) as e:
        bs.outlines()
    assert e.value.code == -3                               # nothing submitted
    bs.apply([bg])
    assert bs.outlines() == ([], [])
