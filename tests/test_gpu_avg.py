"""GPU background generation (tb_avg_*) vs the oracle restatement of AveragingAccumulator; bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["mean", "mode", "max", "min"])
def test_average_vs_oracle(mode):
    import trex_b200
    from oracle import seg
    rng = np.random.default_rng(8)
    h, w, n = 120, 200, 37
    base = rng.integers(90, 160, (h, w))
    fr = np.clip(base[None] + rng.integers(-6, 7, (n, h, w)), 0, 255).astype(np.uint8)
    fr[5:9, 40:80, 60:100] = 10                      # a fish passing through
    acc = trex_b200.AveragingAccumulator(w, h, mode)
    acc.add(fr[:20]); acc.add(fr[20]); acc.add(fr[21:])
    got = acc.finalize()
    assert np.array_equal(got, seg.average(fr, mode))


def test_average_feeds_segmentation():
    """generate_average -> set_background -> apply, end to end on the GPU."""
    import trex_b200
    from oracle import seg
    from trex_b200.synthetic import BlobWorld
    world = BlobWorld(h=272, w=480, n_blobs=10, seed=2, margin=30)
    frames = world.frames(24)
    acc = trex_b200.AveragingAccumulator(480, 272, "mode")
    acc.add(frames)
    bg = acc.finalize()
    assert np.array_equal(bg, seg.average(frames, "mode"))
    bs = trex_b200.BackgroundSubtraction(bg, max_batch=4)
    got = bs.apply(frames[:4])
    P = seg.Params(detect_threshold=15, detect_size_filter=[(10.0, 100000.0)])
    for f in range(4):
        assert [(b.lines.tobytes(), b.pixels.tobytes()) for b in got[f]] == seg.segment_frame(frames[f], bg, P).as_list()
    with pytest.raises(trex_b200.TrexB200Error):
        trex_b200.AveragingAccumulator(480, 272, "mean").finalize()      # no samples


def test_average_golden_from_test_pv():
    """GPU mode-averaging reproduces the background TRex stored in videos/test.pv (committed window)."""
    import os
    import trex_b200
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "avg_golden.npz"))
    fr = g["frames"]
    acc = trex_b200.AveragingAccumulator(fr.shape[2], fr.shape[1], "mode")
    acc.add(fr)
    assert np.array_equal(acc.finalize(), g["expected"])
