"""GPU parity tests of the colour path (BGR / BGRA frames; meta_encoding gray and rgb8): CUDA through the C ABI vs the
CPU oracle, bit-exact.  Reference: T/python/BackgroundSubtraction.cpp:151-188, C/processing/RawProcessing.cpp:355-358,
557-593, C/processing/Source.cpp:200-231, T/tracking/FilterCache.cpp:158-235."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _as_list(blobs):
    return [(b.lines.tobytes(), b.pixels.tobytes()) for b in blobs]


def _world(h, w, n, seed, channels):
    from trex_b200.synthetic import BlobWorld, to_color
    world = BlobWorld(h=h, w=w, n_blobs=n, seed=seed, margin=30)
    gray = world.frames(3)
    frames = to_color(gray, seed=seed, channels=channels)
    bg3 = to_color(world.bg, seed=seed + 1, channels=3)
    # adversarial pixels inside / outside blobs: grey value 0 with a non-zero channel, all-zero pixels, saturated pixels
    rng = np.random.default_rng(seed)
    for f in range(len(frames)):
        for _ in range(200):
            y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
            frames[f, y, x, :3] = [(4, 0, 0), (0, 0, 0), (255, 255, 255), (0, 1, 0)][int(rng.integers(0, 4))]
    return frames, bg3


def _mk(bg, channels, encoding, max_batch=4, max_individuals=0, max_runs_per_frame=0, **kw):
    import trex_b200
    s = trex_b200.DetectSettings(meta_encoding=encoding, **kw)
    return trex_b200.BackgroundSubtraction(bg, settings=s, max_batch=max_batch, max_individuals=max_individuals, channels=channels,
                                           max_runs_per_frame=max_runs_per_frame)


def _params(**kw):
    from oracle import seg
    keys = {k: v for k, v in kw.items() if k in seg.Params.__dataclass_fields__}
    return seg.Params(**keys)


@pytest.mark.parametrize("channels", [3, 4])
@pytest.mark.parametrize("size", [(272, 480), (1080, 1920), (123, 250)])
def test_gray_encoding_from_colour_frames(channels, size):
    """cvtColor(BGR[A]2GRAY) fused into K1 (aligned widths) or through the grey plane (250 is not a multiple of 16)."""
    from oracle import seg
    h, w = size
    frames, bg3 = _world(h, w, 30, 7, channels)
    bg = seg.bgr2gray(bg3)
    kw = dict(detect_threshold=15, detect_size_filter=[(1, 100000)])
    bs = _mk(bg, channels, "gray", max_individuals=40, **kw)
    got = bs.apply(frames)
    crops, _ = bs.crops()
    n = 0
    for f in range(len(frames)):
        ref = seg.segment_frame_color(frames[f], bg, _params(**kw), encoding=seg.ENC_GRAY)
        assert _as_list(got[f]) == ref.as_list(), f
        assert np.array_equal(bs.debug_binary(frames[f]), seg.generate_binary_color(frames[f], bg, _params(**kw))[0])
        for k in range(min(len(ref), 40)):
            assert np.array_equal(crops[n], seg.crop_blob(*ref.blob(k), bg, seg.DIFF_ABSOLUTE)), (f, k)
            n += 1
    assert n == len(crops) and n > 0


@pytest.mark.parametrize("channels", [3, 4])
@pytest.mark.parametrize("size", [(272, 480), (1080, 1920), (123, 250)])
def test_rgb8_encoding(channels, size):
    """3 bytes per blob pixel, foreground = grey difference > T and any of B,G,R != 0, 80x80x3 crops."""
    from oracle import seg
    h, w = size
    frames, bg3 = _world(h, w, 30, 11, channels)
    kw = dict(detect_threshold=15, detect_size_filter=[(1, 100000)])
    bs = _mk(bg3, channels, "rgb8", max_individuals=40, **kw)
    got = bs.apply(frames)
    crops, _ = bs.crops()
    assert crops.shape[1:] == (80, 80, 3)
    n = 0
    for f in range(len(frames)):
        ref = seg.segment_frame_color(frames[f], bg3, _params(**kw), encoding=seg.ENC_RGB8)
        assert len(ref) > 0
        assert _as_list(got[f]) == ref.as_list(), f
        assert np.array_equal(bs.debug_binary(frames[f]), seg.generate_binary_color(frames[f], bg3, _params(**kw), encoding=seg.ENC_RGB8)[0])
        for k in range(min(len(ref), 40)):
            assert np.array_equal(crops[n], seg.crop_blob_rgb(*ref.blob(k), bg3, seg.DIFF_ABSOLUTE)), (f, k)
            n += 1
    assert n == len(crops) and n > 0


@pytest.mark.parametrize("encoding", ["gray", "rgb8"])
@pytest.mark.parametrize("kw", [
    dict(detect_threshold=200), dict(detect_threshold=-15), dict(detect_threshold=12, threshold_maximum=60),
    dict(detect_threshold=15, detect_threshold_is_absolute=False), dict(detect_threshold=15, image_invert=True),
    dict(detect_threshold=15, use_closing=True, closing_size=2), dict(detect_threshold=15, dilation_size=3),
])
def test_colour_non_default_settings(encoding, kw):
    """Everything but the default settings runs on the grey plane (plus the non-zero plane for rgb8)."""
    from oracle import seg
    frames, bg3 = _world(144, 256, 12, 3, 3)
    enc = seg.ENC_RGB8 if encoding == "rgb8" else seg.ENC_GRAY
    bg = bg3 if enc else seg.bgr2gray(bg3)
    kw = dict(kw, detect_size_filter=[(1, 1000000)])
    cap = dict(max_runs_per_frame=144 * 256 // 2 + 16, max_pixels_per_frame=144 * 256 * 3)
    import trex_b200
    bs = trex_b200.BackgroundSubtraction(bg, settings=trex_b200.DetectSettings(meta_encoding=encoding, **kw), max_batch=4, channels=3, **cap)
    got = bs.apply(frames)
    for f in range(len(frames)):
        ref = seg.segment_frame_color(frames[f], bg, _params(**kw), encoding=enc)
        assert _as_list(got[f]) == ref.as_list(), f
        assert np.array_equal(bs.debug_binary(frames[f]), seg.generate_binary_color(frames[f], bg, _params(**kw), encoding=enc)[0])


def test_color_channel_selects_a_plane():
    """color_channel (BackgroundSubtraction.cpp:161-173): the plane replaces cvtColor under gray encoding."""
    from oracle import seg
    frames, bg3 = _world(144, 256, 12, 5, 3)
    bg = bg3[..., 1].copy()
    kw = dict(detect_threshold=15, detect_size_filter=[(1, 100000)])
    bs = _mk(bg, 3, "gray", color_channel=1, **kw)
    got = bs.apply(frames)
    for f in range(len(frames)):
        ref = seg.segment_frame(frames[f][..., 1].copy(), bg, _params(**kw))
        assert _as_list(got[f]) == ref.as_list(), f


def test_colour_errors():
    import trex_b200
    from trex_b200._capi import TrexB200Error
    bg = np.full((64, 64), 100, np.uint8)
    with pytest.raises(TrexB200Error):       # rgb8 needs colour frames (BackgroundSubtraction.cpp:177-181)
        trex_b200.BackgroundSubtraction(bg, settings=trex_b200.DetectSettings(meta_encoding="rgb8"), channels=1)
    with pytest.raises(TrexB200Error):       # rgb8 needs a 3-channel background (RawProcessing.cpp:343)
        trex_b200.BackgroundSubtraction(bg, settings=trex_b200.DetectSettings(meta_encoding="rgb8"), channels=3)
    with pytest.raises(TrexB200Error):       # r3g3b2 needs colour frames too (BackgroundSubtraction.cpp:151-158)
        trex_b200.BackgroundSubtraction(bg, settings=trex_b200.DetectSettings(meta_encoding="r3g3b2"), channels=1)
    with pytest.raises(TrexB200Error):       # and a 1-channel background of codes
        trex_b200.BackgroundSubtraction(np.zeros((64, 64, 3), np.uint8), settings=trex_b200.DetectSettings(meta_encoding="r3g3b2"), channels=3)
    bs = trex_b200.BackgroundSubtraction(bg, channels=3)
    with pytest.raises(TrexB200Error):       # wrong frame shape
        bs.apply([np.zeros((64, 64), np.uint8)])


def test_rethreshold_on_colour_frames_gray_encoding():
    """Tracker-side re-threshold of blobs detected on BGR frames (gray encoding): equals the 1-channel path on cvtColor."""
    import trex_b200
    from oracle import seg
    frames, bg3 = _world(272, 480, 20, 9, 3)
    bg = seg.bgr2gray(bg3)
    kw = dict(detect_threshold=15, detect_size_filter=[(1, 100000)])
    det = _mk(bg, 3, "gray", **kw)
    trk = trex_b200.BackgroundSubtraction(bg, settings=trex_b200.DetectSettings(detect_threshold=40, detect_size_filter=[]), max_batch=4)
    det.apply(frames)
    got = det.rethreshold(trk)
    for f in range(len(frames)):
        ref = seg.rethreshold(seg.segment_frame_color(frames[f], bg, _params(**kw)), bg, 40, seg.DIFF_ABSOLUTE)
        assert set(_as_list(got[f])) == ref.as_set(), f


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "fp16"])
def test_rgb8_chain_frames_to_identities(precision):
    """BGRA frames -> rgb8 blobs -> 80x80x3 crops -> V118_3(channels=3), chained on the device like bench.py does."""
    import torch
    import trex_b200
    from oracle import seg, vi
    frames, bg3 = _world(272, 480, 12, 21, 4)
    kw = dict(detect_threshold=15, detect_size_filter=[(10, 100000)])
    bs = _mk(bg3, 4, "rgb8", max_individuals=16, **kw)
    M = 12
    sd = vi.scale_for_u8_inputs(vi.init_state_dict(M, 3, 80, 80, seed=0))
    net = trex_b200.VINetwork(M, channels=3, max_images=64, precision=precision)
    net.load_weights(sd)
    bs.apply(frames)
    crops, _ = bs.crops()
    exp = []
    for f in range(len(frames)):
        ref = seg.segment_frame_color(frames[f], bg3, _params(**kw), encoding=seg.ENC_RGB8)
        exp += [seg.crop_blob_rgb(*ref.blob(k), bg3, seg.DIFF_ABSOLUTE) for k in range(min(len(ref), 16))]
    exp = np.stack(exp)
    assert np.array_equal(crops, exp)
    # device chain: crops never leave HBM
    crops_p, ncrops_p, _, _, _ = bs.device_results()
    dev = torch.device("cuda", 0)
    probs = torch.zeros((64, M), dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream(dev)
    bs.apply_device(torch.from_numpy(frames).to(dev).data_ptr(), len(frames), stream.cuda_stream)
    net.predict_device(crops_p, 64, ncrops_p, probs.data_ptr(), 0, stream.cuda_stream)
    net.wait()
    got = probs.cpu().numpy()[:len(exp)]
    assert np.abs(got - vi.predict(sd, exp)).max() < 1e-3


@pytest.mark.parametrize("method", ["absolute", "sign"])
def test_rethreshold_rgb8_blobs(method):
    """Tracker-side re-threshold of rgb8 blobs: cmn::bgr2gray(pixel) vs the background's grey image, comparison >=,
    B,G,R bytes kept (oracle pinned on test_pixels.cpp:1073-1166, 1289-1379)."""
    import trex_b200
    from oracle import seg
    frames, bg3 = _world(272, 480, 20, 13, 3)
    kw = dict(detect_threshold=15, detect_size_filter=[(1, 100000)])
    det = _mk(bg3, 3, "rgb8", **kw)
    trk_settings = trex_b200.DetectSettings(meta_encoding="rgb8", detect_threshold=35, detect_size_filter=[],
                                            detect_threshold_is_absolute=(method == "absolute"))
    trk = trex_b200.BackgroundSubtraction(bg3, settings=trk_settings, max_batch=4, channels=3)
    det.apply(frames)
    got = det.rethreshold(trk)
    bg_gray = seg.bgr2gray(bg3)
    m = seg.DIFF_ABSOLUTE if method == "absolute" else seg.DIFF_SIGN
    total = 0
    for f in range(len(frames)):
        ref = seg.rethreshold(seg.segment_frame_color(frames[f], bg3, _params(**kw), encoding=seg.ENC_RGB8), bg_gray, 35, m, rgb=True)
        assert set(_as_list(got[f])) == ref.as_set(), f
        total += len(ref)
    assert total > 0


@pytest.mark.parametrize("channels", [3, 4])
@pytest.mark.parametrize("size,crop_method", [((272, 480), "absolute"), ((1080, 1920), "absolute"), ((123, 250), "sign"), ((272, 480), "none")])
def test_r3g3b2_encoding(channels, size, crop_method):
    """meta_encoding r3g3b2: frames become 1-byte codes (convert_to_r3g3b2), the 1-channel path runs on the codes against a
    background of codes; one code byte per blob pixel; crops are B,G,R renderings (r3g3b2_to_vec) differenced per channel."""
    from oracle import seg
    h, w = size
    frames, bg3 = _world(h, w, 30, 13, channels)
    bg = seg.convert_to_r3g3b2(bg3)
    method = {"none": seg.DIFF_NONE, "absolute": seg.DIFF_ABSOLUTE, "sign": seg.DIFF_SIGN}[crop_method]
    kw = dict(detect_threshold=20, detect_size_filter=[(1, 100000)],
              track_background_subtraction=crop_method != "none", track_threshold_is_absolute=crop_method != "sign")
    # the codes of a noisy frame flicker around the quantisation steps: ~20k small blobs per 1080p frame -> large run capacity
    bs = _mk(bg, channels, "r3g3b2", max_individuals=40, max_runs_per_frame=1 << 17, **kw)
    got = bs.apply(frames)
    crops, _ = bs.crops()
    assert crops.shape[1:] == (80, 80, 3)
    n = 0
    for f in range(len(frames)):
        ref = seg.segment_frame_color(frames[f], bg, _params(**kw), encoding=seg.ENC_R3G3B2)
        assert len(ref) > 0
        assert _as_list(got[f]) == ref.as_list(), f
        assert np.array_equal(bs.debug_binary(frames[f]), seg.generate_binary_color(frames[f], bg, _params(**kw), encoding=seg.ENC_R3G3B2)[0])
        for k in range(min(len(ref), 40)):
            assert np.array_equal(crops[n], seg.crop_blob_r3g3b2(*ref.blob(k), bg, method)), (f, k)
            n += 1
    assert n == len(crops) and n > 0


def test_r3g3b2_chain_to_identities():
    """r3g3b2 crops (80x80x3) feed the 3-channel V118_3 like rgb8 crops do."""
    import trex_b200
    from oracle import seg, vi
    frames, bg3 = _world(272, 480, 12, 5, 3)
    bg = seg.convert_to_r3g3b2(bg3)
    kw = dict(detect_threshold=20, detect_size_filter=[(20, 100000)])
    bs = _mk(bg, 3, "r3g3b2", max_individuals=16, **kw)
    bs.apply(frames)
    crops, _ = bs.crops()
    assert len(crops) > 0
    sd = vi.scale_for_u8_inputs(vi.init_state_dict(10, 3, 80, 80, seed=0))
    net = trex_b200.VINetwork(10, channels=3, max_images=64)
    net.load_weights(sd)
    probs = net.probabilities(crops)
    assert np.abs(probs - vi.predict(sd, crops)).max() < 1e-3


def test_batch_arena_overflow_is_reported_not_corrupting():
    """More blobs in a batch than the blob / line arenas hold: tb_seg_wait reports TB_ERR_CAPACITY, the frames before the
    overflow keep their results and the batch totals stay inside the arenas."""
    import trex_b200
    from trex_b200._capi import TrexB200Error
    from oracle import seg
    rng = np.random.default_rng(2)
    h, w = 256, 512
    bg = np.full((h, w), 100, np.uint8)
    frames = np.repeat(bg[None], 4, 0).copy()
    frames[:, ::2, ::2] = 200                                   # 32768 single-pixel blobs per frame
    kw = dict(detect_threshold=15, detect_size_filter=[(1, 100000)])
    bs = trex_b200.BackgroundSubtraction(bg, settings=trex_b200.DetectSettings(**kw), max_batch=4, max_runs_per_frame=40000)
    with pytest.raises(TrexB200Error) as e:
        bs.apply(frames)
    assert e.value.code == -4
    tb, tl, tp, _ = bs.totals()
    assert tb == tl == tp == 32768                              # arenas hold max(4 * 2048, 40000) blobs: one frame fits
    infos = [bs.frame_info(i) for i in range(4)]
    assert infos[0].n_blobs == 32768 and infos[0].status == 0
    assert all(i.n_blobs == 0 and (i.status & 8) for i in infos[1:])
    ref = seg.segment_frame(frames[0], bg, _params(**kw))
    assert _as_list(bs.result(0)) == ref.as_list()


def test_rgb8_size_filter_counts_payload_bytes():
    """detect_size_filter on rgb8 blobs compares 3 * pixels (pixels->size(), BackgroundSubtraction.cpp:247-259): blobs whose
    pixel count straddles the bounds are kept / dropped like the oracle does (which pins the rule on a hand-made frame)."""
    from oracle import seg
    bg = np.full((48, 64, 3), 128, np.uint8)
    fr = np.repeat(bg[None], 2, 0).copy()
    fr[0, 5, 8:12] = (20, 30, 40); fr[0, 20, 10:44] = (20, 30, 40); fr[0, 30, 10:20] = (20, 30, 40)
    fr[1, 7, 3:6] = (20, 30, 40); fr[1, 9, 3:7] = (200, 30, 40); fr[1, 40, 0:33] = (20, 30, 40); fr[1, 42, 0:34] = (20, 30, 40)
    kw = dict(detect_threshold=15, detect_size_filter=[(10, 100)])
    bs = _mk(bg, 3, "rgb8", **kw)
    got = bs.apply(fr)
    for f in range(2):
        ref = seg.segment_frame_color(fr[f], bg, _params(**kw), encoding=seg.ENC_RGB8)
        assert _as_list(got[f]) == ref.as_list(), f
    assert sorted(b.num_pixels for b in got[0]) == [4, 10] and sorted(b.num_pixels for b in got[1]) == [4, 33]
    gs = _mk(seg.bgr2gray(bg), 3, "gray", **kw)
    assert sorted(b.num_pixels for b in gs.apply(fr)[0]) == [10, 34]


def test_recount_rgb8_blobs():
    """rgb8 blobs are recounted through cmn::bgr2gray of each pixel against the background's grey image (diffable_pixel_value<rgb8 -> gray>)."""
    import trex_b200
    from oracle import seg as oseg
    rng = np.random.default_rng(12)
    H, W = 96, 160
    bg3 = rng.integers(120, 200, (H, W, 3)).astype(np.uint8)
    fr = bg3.copy()
    fr[20:50, 30:90] = rng.integers(0, 110, (30, 60, 3)).astype(np.uint8)
    fr[60:80, 100:140] = rng.integers(0, 255, (20, 40, 3)).astype(np.uint8)
    s = trex_b200.DetectSettings(meta_encoding="rgb8", detect_size_filter=[])
    bs = trex_b200.BackgroundSubtraction(bg3, settings=s, max_batch=1, channels=3)
    got = bs.apply([fr])[0]
    assert len(got) >= 2
    bg_gray = oseg.bgr2gray(bg3)
    for T in (10, 50):
        rc = bs.recount(T)
        exp = np.array([oseg.blob_recount(b.lines, b.pixels, bg_gray, T, oseg.DIFF_ABSOLUTE, channels=3) for b in got], np.float32)
        assert np.array_equal(rc, exp), T
