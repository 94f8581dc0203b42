"""N>1 host logic on CPU: world_size-2 gloo processes shard frames, all-gather the fixed-stride blob
metadata and every rank reassembles all frames in order."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from trex_b200 import sharding
from trex_b200.background_subtraction import INFO_DTYPE, REC_DTYPE

B, KMAX, ROUNDS = 4, 8, 3


def _fake_meta(rank, world, rnd):
    """Deterministic metadata a rank would produce for its frames of a round."""
    lo, hi = sharding.frame_range(rnd, rank, world, B)
    infos = np.zeros(B, INFO_DTYPE)
    recs = np.zeros(B * KMAX, REC_DTYPE)
    k = 0
    for i, f in enumerate(range(lo, hi)):
        n = (f * 7 + 3) % (KMAX + 1)
        infos[i]["blob_begin"], infos[i]["n_blobs"] = k, n
        for j in range(n):
            recs[k]["bid"] = f * 1000 + j
            recs[k]["frame"] = i
            recs[k]["n_pixels"] = 10 + j
            k += 1
    return infos, recs


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        for rnd in range(ROUNDS):
            infos, recs = _fake_meta(rank, world, rnd)
            n_all = int(infos["n_blobs"].sum())
            top_id = np.zeros(B * KMAX, np.uint32); top_p = np.zeros(B * KMAX, np.float32)
            top_id[:n_all] = recs["bid"][:n_all] % 97; top_p[:n_all] = 1.0 / (1 + recs["n_pixels"][:n_all])
            local = sharding.pack_metadata(infos, recs, B, KMAX, top_id, top_p)
            assert local.numel() == sharding.meta_bytes(B, KMAX)
            gathered = sharding.all_gather_metadata(local)
            frames = sharding.unpack_round(gathered, rnd, B, KMAX, with_identity=True)
            exp_frames = list(range(rnd * world * B, (rnd + 1) * world * B))
            ok &= list(frames) == exp_frames
            for f, (info, r, trunc, ids, ps) in frames.items():
                n = (f * 7 + 3) % (KMAX + 1)
                ok &= int(info["n_blobs"]) == n and not trunc
                ok &= [int(x) for x in r["bid"]] == [f * 1000 + j for j in range(n)]
                ok &= [int(x) for x in ids] == [(f * 1000 + j) % 97 for j in range(n)]           # identities travel with their blobs
                ok &= np.allclose(ps, [1.0 / (11 + j) for j in range(n)])
                rr, rk, idx = sharding.owner_of(f, world, B)
                ok &= rr == rnd and sharding.frame_range(rr, rk, world, B)[0] + idx == f
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_two_rank_metadata_allgather():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_frame_partition_covers_everything_once():
    world, batch = 8, 64
    seen = []
    for rnd in range(3):
        for rank in range(world):
            lo, hi = sharding.frame_range(rnd, rank, world, batch)
            seen.extend(range(lo, hi))
    assert seen == list(range(3 * world * batch))


def test_meta_layout_mirror():
    """Host mirror of tb_meta_layout: 32-byte aligned sections, records last (the gathered prefix ends inside the record array)."""
    lay = sharding.MetaLayout.make(256, 128)
    assert lay.off_top_id == 256 * 32 and lay.off_top_p == lay.off_top_id + 256 * 128 * 4
    assert lay.off_recs == lay.off_top_p + 256 * 128 * 4 and lay.gather_bytes == lay.off_recs + 256 * 128 * 32
    odd = sharding.MetaLayout.make(3, 5)
    assert odd.off_top_id % 32 == 0 and odd.off_top_p % 32 == 0 and odd.off_recs % 32 == 0
    assert odd.off_top_p >= odd.off_top_id + 60 and odd.off_recs >= odd.off_top_p + 60


def test_bench_balanced_batches():
    """bench.py's end-to-end frame shares: equal unless the ranks' concurrent H2D rates differ by more than 10 %, then proportional to the rate."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.balanced_batches(256, [55.0]) == [256]
    assert bench.balanced_batches(256, [52.4, 55.2, 52.5, 55.1]) == [256] * 4
    got = bench.balanced_batches(256, [23.24, 35.71, 23.37, 35.62, 23.32, 35.54, 23.3, 35.49])
    assert got[1] == 256 and got[0] == 167 and all(1 <= g <= 256 for g in got)
    assert abs(sum(got) * 2.0736e6 / 1e9 / max(g * 2.0736e6 / 1e9 / r for g, r in zip(got, [23.24, 35.71, 23.37, 35.62, 23.32, 35.54, 23.3, 35.49])) - 235.6) < 3   # all ranks finish together
    assert bench.balanced_batches(64, [0.5, 55.0]) == [1, 64]
