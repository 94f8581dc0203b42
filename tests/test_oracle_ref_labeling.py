"""Pins the oracle's run-length connected-component labeling (oracle/trex_oracle.c: extract_lines, merge_lines with the Brototype survivor rule,
run_fast's emission order "current source" = ORDER_REF_LAZY; SURVEY.md s8 rows a-4 ... a-6) on the REFERENCE'S OWN CODE:
commons/common/processing/{CPULabeling,Brototype,Source,DLList,ListCache}.cpp, compiled unmodified from the reference checkout (oracle/build_ref.py;
cv::Mat replaced by a plain byte image).  Blob for blob IN THE REFERENCE'S ORDER: run lists, pixel bytes (1 and 3 per pixel), the is_rgb flag; through
both entries -- run(image, cache) (detection) and run(lines, pixels, cache, channels) (pixel::threshold_blob).  The `videos/test.pv` fixture
(tests/test_oracle_golden.py) pins the same code path on real data; this one adds adversarial images, colour, very wide runs and the emission order
of the current source.  Runs wherever oracle/_ref/libref_posture.so exists or can be built; skipped otherwise."""
import ctypes as C

import numpy as np
import pytest

from oracle import build_ref, seg


@pytest.fixture(scope="module")
def ref():
    path = build_ref.build()
    if path is None:
        pytest.skip("no reference checkout and no prebuilt oracle/_ref/libref_posture.so")
    lib = C.CDLL(path)
    lib.ref_label_image.restype = C.c_int64
    lib.ref_label_lines.restype = C.c_int64
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _unpack(k, lines, px, lo, po, fl):
    assert k >= 0, k
    return [(lines[lo[i]:lo[i + 1], :3].copy(), px[po[i]:po[i + 1]].copy(), int(fl[i])) for i in range(k)]


def ref_label_image(ref, img):
    h, w = img.shape[:2]
    ch = 1 if img.ndim == 2 else img.shape[2]
    npx = int((img.reshape(h, w, -1).max(2) > 0).sum())
    lines = np.zeros((npx + 8, 4), np.uint16); px = np.zeros(npx * ch + 8, np.uint8)
    lo = np.zeros(npx + 9, np.int64); po = np.zeros(npx + 9, np.int64); fl = np.zeros(npx + 8, np.uint8)
    k = ref.ref_label_image(_p(np.ascontiguousarray(img)), h, w, ch, _p(lines), C.c_int64(len(lines)), _p(px), C.c_int64(len(px)), _p(lo), _p(po), _p(fl), C.c_int64(len(fl)))
    return _unpack(k, lines, px, lo, po, fl)


def ref_label_lines(ref, runs, pixels, ch):
    raw = np.zeros((len(runs), 4), np.uint16); raw[:, :3] = runs
    n = len(runs)
    lines = np.zeros((n + 8, 4), np.uint16); px = np.zeros(len(pixels) + 8, np.uint8)
    lo = np.zeros(n + 9, np.int64); po = np.zeros(n + 9, np.int64); fl = np.zeros(n + 8, np.uint8)
    pixels = np.ascontiguousarray(pixels, np.uint8)
    k = ref.ref_label_lines(_p(raw), C.c_int64(n), _p(pixels), C.c_int64(len(pixels)), ch, _p(lines), C.c_int64(len(lines)), _p(px), C.c_int64(len(px)),
                            _p(lo), _p(po), _p(fl), C.c_int64(len(fl)))
    return _unpack(k, lines, px, lo, po, fl)


def oracle_blobs(B):
    out = []
    for b in range(len(B)):
        l, p = B.blob(b)
        out.append((np.stack([l["x0"], l["x1"], l["y"]], 1).astype(np.uint16), np.asarray(p)))
    return out


def same(mine, want):
    return len(mine) == len(want) and all(np.array_equal(a[0], r[0]) and np.array_equal(a[1], r[1]) for a, r in zip(mine, want))


def gray_images():
    rng = np.random.default_rng(2)
    out = [((rng.random((60, 90)) < d) * rng.integers(1, 255, (60, 90))).astype(np.uint8) for d in (0.2, 0.3, 0.45, 0.5, 0.6, 0.7, 0.9)]
    comb = np.zeros((40, 64), np.uint8); comb[::2, :] = 200; comb[:, 0] = 200; comb[5, 10:20] = 0        # a comb: many runs merging into one blob late
    out.append(comb)
    spiral = np.zeros((41, 41), np.uint8)
    for r in range(0, 20, 2):
        spiral[r, r:41 - r] = 255; spiral[40 - r, r:41 - r] = 255; spiral[r:41 - r, r] = 255; spiral[r + 2:41 - r, 40 - r] = 255
    out.append(spiral)
    checker = (np.indices((30, 30)).sum(0) % 2 * 255).astype(np.uint8)                                   # 8-connectivity: one blob through the diagonals
    out.append(checker)
    out.append(np.zeros((16, 16), np.uint8))                                                              # empty
    wide = np.zeros((6, 9000), np.uint8); wide[1] = 7; wide[2, 100:8900] = 9; wide[4, ::3] = 1            # runs far longer than any packed line field
    out.append(wide)
    return out


def test_gray_images_blob_for_blob_in_the_references_order(ref):
    n = 0
    for img in gray_images():
        want = ref_label_image(ref, img)
        mine = oracle_blobs(seg.label_image(img, seg.ORDER_REF_LAZY))
        assert same(mine, [(r[0], r[1]) for r in want]), img.shape
        assert all(r[2] == 0 for r in want)
        # the canonical order the GPU emits is a permutation of it
        canon = oracle_blobs(seg.label_image(img, seg.ORDER_CANONICAL))
        key = lambda t: (t[0].tobytes(), t[1].tobytes())
        assert sorted(map(key, canon)) == sorted(map(key, mine))
        n += len(want)
    assert n > 1000


def test_lines_entry_equals_image_entry(ref):
    """run(lines, pixels, cache, channels) -- what pixel::threshold_blob calls -- on all runs of an image gives the blobs of run(image, cache)."""
    for img in gray_images()[:6]:
        want = ref_label_image(ref, img)
        runs, px = [], []
        for y in range(img.shape[0]):
            row = np.flatnonzero(np.diff(np.concatenate([[0], (img[y] > 0).astype(np.int8), [0]])))
            for x0, x1 in zip(row[::2], row[1::2]):
                runs.append((x0, x1 - 1, y)); px.append(img[y, x0:x1])
        got = ref_label_lines(ref, np.array(runs, np.uint16).reshape(-1, 3), np.concatenate(px) if px else np.zeros(0, np.uint8), 1)
        assert same([(g[0], g[1]) for g in got], [(w[0], w[1]) for w in want])


def test_three_channel_image_keeps_the_colour_bytes(ref):
    rng = np.random.default_rng(4)
    for d in (0.3, 0.5, 0.7):
        mask = rng.random((50, 70)) < d
        img = (mask[..., None] * rng.integers(40, 255, (50, 70, 3))).astype(np.uint8)      # every foreground pixel well above the grey threshold
        want = ref_label_image(ref, img)
        P = seg.Params(detect_threshold=15, enable_difference=True, detect_size_filter=[])
        mine = oracle_blobs(seg.segment_frame_color(img, np.zeros_like(img), P, seg.ENC_RGB8, -1, seg.ORDER_REF_LAZY))
        assert same(mine, [(r[0], r[1]) for r in want])
        assert all(r[2] == 1 << 5 for r in want)                                            # pv::Blob::Flags::is_rgb
        assert all(len(r[1]) == 3 * int((r[0][:, 1].astype(int) - r[0][:, 0] + 1).sum()) for r in want)


def ref_threshold_blob(ref, runs, pixels, ch, diff, threshold):
    raw = np.zeros((len(runs), 4), np.uint16); raw[:, :3] = runs
    n = len(runs)
    cap = len(diff) + 8
    lines = np.zeros((cap, 4), np.uint16); px = np.zeros(len(pixels) + 8, np.uint8)
    lo = np.zeros(cap + 1, np.int64); po = np.zeros(cap + 1, np.int64); fl = np.zeros(cap, np.uint8)
    pixels = np.ascontiguousarray(pixels, np.uint8); diff = np.ascontiguousarray(diff, np.uint8)
    ref.ref_threshold_blob_cache.restype = C.c_int64
    k = ref.ref_threshold_blob_cache(_p(raw), C.c_int64(n), _p(pixels), C.c_int64(len(pixels)), ch, _p(diff), int(threshold), _p(lines), C.c_int64(len(lines)), _p(px),
                                     C.c_int64(len(px)), _p(lo), _p(po), _p(fl), C.c_int64(len(fl)))
    return _unpack(k, lines, px, lo, po, fl)


@pytest.mark.parametrize("method", [seg.DIFF_ABSOLUTE, seg.DIFF_SIGN, seg.DIFF_NONE])
def test_tracker_rethreshold_against_the_compiled_threshold_blob(ref, method):
    """pixel::threshold_blob (the entry the tracker calls) on every blob of noisy frames: the reference's compiled run cutting + relabeling + its
    `pixels->size() > 1` rule (a lone grey pixel is dropped) against oracle rethreshold().  The reference is driven through its difference-cache
    overload with the difference values of Background.h:231-294 (none: v, absolute: |bg - v|, sign: max(0, bg - v)) computed here; the set of
    sub-blobs of every parent must match (the oracle emits them in canonical order, the reference in merge order)."""
    rng = np.random.default_rng(9)
    n_sub = n_single = n_row0 = 0
    for trial in range(4):
        h, w = 48, 64
        bg = rng.integers(90, 160, (h, w)).astype(np.uint8)
        frame = bg.copy()
        mask = rng.random((h, w)) < 0.45
        frame[mask] = np.clip(bg[mask].astype(int) + rng.integers(-120, 120, int(mask.sum())), 0, 255).astype(np.uint8)
        P = seg.Params(detect_threshold=10, detect_size_filter=[])
        parents = seg.segment_frame(frame, bg, P)
        T = 45
        mine = seg.rethreshold(parents, bg, T, method)
        mine_set = sorted((l.tobytes(), p.tobytes()) for l, p in oracle_blobs(mine))
        n_mine = len(mine_set)
        want_set = []
        for b in range(len(parents)):
            l, p = parents.blob(b)
            runs = np.stack([l["x0"], l["x1"], l["y"]], 1).astype(np.uint16)
            vals = np.asarray(p).astype(int)
            bgv = np.concatenate([bg[y, x0:x1 + 1] for x0, x1, y in runs]).astype(int)
            diff = {seg.DIFF_NONE: vals, seg.DIFF_ABSOLUTE: np.abs(bgv - vals), seg.DIFF_SIGN: np.maximum(0, bgv - vals)}[method]
            got = ref_threshold_blob(ref, runs, np.asarray(p), 1, diff.astype(np.uint8), T)
            if len(got) == 0 and (diff >= T).any():
                one = seg.Blobs(l.copy(), np.asarray(p).copy(), np.array([0, len(l)], np.int64), np.array([0, len(p)], np.int64))
                mine_one = oracle_blobs(seg.rethreshold(one, bg, T, method))
                if mine_one:
                    # the one deliberate deviation: when every surviving run lies in image row 0, Source::RowRef::from_index(0) takes the
                    # "beyond the value ranges" exit (Source.h:150-156) and the reference returns NO blobs; the oracle and the GPU return them
                    ys = np.repeat(runs[:, 2], runs[:, 1].astype(int) - runs[:, 0] + 1)
                    assert set(int(v) for v in ys[diff >= T]) == {0}
                    for ml, mp in mine_one:
                        mine_set.remove((ml.tobytes(), mp.tobytes()))
                    n_row0 += 1
            for sl, sp, _ in got:
                want_set.append((sl.tobytes(), sp.tobytes()))
                n_sub += 1
        assert mine_set == sorted(want_set), (trial, method, len(mine_set), len(want_set))
        # the rule itself: with every sub-blob kept, the oracle has exactly the lone pixels in addition
        all_sub = oracle_blobs(seg.rethreshold(parents, bg, T, method, keep_single=True))
        lone = [s for s in all_sub if len(s[1]) == 1]
        n_single += len(lone)
        assert len(all_sub) == n_mine + len(lone)
    assert n_sub > 200 and n_single > 20


def test_reference_returns_nothing_when_every_run_is_in_row_zero(ref):
    """Documents the deviation the tests above step around: Source::row(0) on a source whose only populated row is y = 0 is invalid (upper_bound
    runs off the end before the `*(it - 1) == y` check, Source.h:150-160), so CPULabeling::run returns no blobs -- through both entries.  The oracle
    and the GPU return the blobs of such a frame; deliberately not reproduced."""
    img = np.zeros((4, 20), np.uint8); img[0, 3:6] = 9; img[0, 14:16] = 7
    assert ref_label_image(ref, img) == []
    assert len(seg.label_image(img, seg.ORDER_REF_LAZY)) == 2
    img[2, 8] = 5                                                   # any other populated row, and row 0 is seen
    assert len(ref_label_image(ref, img)) == 3
    assert ref_label_lines(ref, np.array([[14, 15, 0]], np.uint16), np.array([1, 2], np.uint8), 1) == []
    assert len(ref_label_lines(ref, np.array([[14, 15, 1]], np.uint16), np.array([1, 2], np.uint8), 1)) == 1


def test_blob_id_against_the_references_own_bid(ref):
    """pv::blob_bid (processing/BlobIdentity.cpp) over pv::bid::from_data (misc/bid.h:87-94), both compiled from the checkout: the id of a blob from its first run
    and its run count -- which passes through a uint8_t parameter before the % 64 (so 256 runs count as 0, 300 as 44).  Against seg.blob_id (= what K3 writes
    into tb_blob_rec.bid) for labelled blobs and for constructed run lists with 1 ... 1000 runs and coordinates up to 8191."""
    ref.ref_blob_bid.restype = C.c_uint32
    rng = np.random.default_rng(12)
    n = 0
    for img in gray_images():
        for runs, _, _ in ref_label_image(ref, img):
            if runs[0, 1] >= 8192:                       # from_data asserts 13-bit coordinates (bid.h:88-90); the 9000-pixel run of gray_images() is beyond them
                continue
            raw = np.zeros((len(runs), 4), np.uint16); raw[:, :3] = runs
            l = np.zeros(len(runs), seg.LINE_DTYPE); l["x0"], l["x1"], l["y"] = runs[:, 0], runs[:, 1], runs[:, 2]
            assert int(ref.ref_blob_bid(_p(raw), C.c_int64(len(raw)))) == seg.blob_id(l)
            n += 1
    for count in (1, 2, 63, 64, 65, 255, 256, 257, 300, 511, 512, 1000):
        for _ in range(4):
            x0 = int(rng.integers(0, 8100)); x1 = x0 + int(rng.integers(0, 8191 - x0)); y = int(rng.integers(0, 8191 - count))
            raw = np.zeros((count, 4), np.uint16); raw[:, 0] = x0; raw[:, 1] = x1; raw[:, 2] = y + np.arange(count)
            l = np.zeros(count, seg.LINE_DTYPE); l["x0"], l["x1"], l["y"] = raw[:, 0], raw[:, 1], raw[:, 2]
            want = (((x0 + x1 + 1) // 2) << 19) | ((y & 0x1FFF) << 6) | ((count & 0xFF) % 64)
            got = int(ref.ref_blob_bid(_p(raw), C.c_int64(count)))
            assert got == seg.blob_id(l) == want, (count, x0, x1, y, got, seg.blob_id(l), want)
            n += 1
    assert n > 300
