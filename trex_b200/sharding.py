"""Frame-batch data parallelism over the GPUs of one box.

Given a fixed background every frame is independent through segmentation, crops and the CNN
(BackgroundSubtraction::apply keeps no cross-frame state, T/python/BackgroundSubtraction.cpp:126-347),
so round k of a G-GPU job gives GPU g the frames  [k*G*B + g*B, k*G*B + (g+1)*B).  The only exchange is
one all-gather per round of the fixed-stride metadata block of every rank so that the host tracker of
every rank sees all frames in order:

    [ tb_frame_info[B] | top-1 identity u32[B*Kmax] | top-1 probability f32[B*Kmax] | tb_blob_rec[B*Kmax] ]

That is exactly the prefix of the device block the library's kernels write in place (tb_seg_metadata /
tb_meta_layout, include/trexb200.h): there is no packing step -- the collective reads the block.
Works with any torch.distributed backend (NCCL on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

from .background_subtraction import INFO_DTYPE, REC_DTYPE


def frame_range(round_idx: int, rank: int, world: int, batch: int):
    """Global frame indices [lo, hi) that `rank` processes in round `round_idx`."""
    lo = (round_idx * world + rank) * batch
    return lo, lo + batch


def owner_of(frame: int, world: int, batch: int):
    """(round, rank, index inside the rank's batch) of a global frame index."""
    chunk = frame // batch
    return chunk // world, chunk % world, frame % batch


def _r32(v: int) -> int:
    return (v + 31) // 32 * 32


@dataclass(frozen=True)
class MetaLayout:
    """Host mirror of tb_meta_layout (offsets in bytes; every section starts on a 32-byte boundary)."""
    batch: int
    kmax: int
    off_infos: int
    off_top_id: int
    off_top_p: int
    off_recs: int
    gather_bytes: int

    @staticmethod
    def make(batch: int, kmax: int) -> "MetaLayout":
        off_top_id = _r32(batch * INFO_DTYPE.itemsize)
        off_top_p = _r32(off_top_id + batch * kmax * 4)
        off_recs = _r32(off_top_p + batch * kmax * 4)
        return MetaLayout(batch, kmax, 0, off_top_id, off_top_p, off_recs, off_recs + batch * kmax * REC_DTYPE.itemsize)

    @staticmethod
    def from_c(m) -> "MetaLayout":
        out = MetaLayout(int(m.batch), int(m.kmax), int(m.off_infos), int(m.off_top_id), int(m.off_top_p), int(m.off_recs), int(m.gather_bytes))
        assert out == MetaLayout.make(out.batch, out.kmax), "tb_meta_layout differs from the host mirror"
        return out


def meta_bytes(batch: int, kmax: int) -> int:
    return MetaLayout.make(batch, kmax).gather_bytes


def pack_metadata(infos: np.ndarray, recs: np.ndarray, batch: int, kmax: int, top_id=None, top_p=None) -> torch.Tensor:
    """HOST-side construction of one rank's block from separate arrays (tests, tools).  The GPU path never calls this:
    the kernels write the block in place."""
    lay = MetaLayout.make(batch, kmax)
    buf = np.zeros(lay.gather_bytes, np.uint8)
    buf[: batch * 32] = np.ascontiguousarray(infos).view(np.uint8)[: batch * 32]
    r = np.ascontiguousarray(recs).view(np.uint8)[: batch * kmax * 32]
    buf[lay.off_recs: lay.off_recs + len(r)] = r
    if top_id is not None:
        buf[lay.off_top_id: lay.off_top_id + batch * kmax * 4] = np.ascontiguousarray(top_id, np.uint32).view(np.uint8)[: batch * kmax * 4]
        buf[lay.off_top_p: lay.off_top_p + batch * kmax * 4] = np.ascontiguousarray(top_p, np.float32).view(np.uint8)[: batch * kmax * 4]
    return torch.from_numpy(buf)


def all_gather_metadata(local: torch.Tensor, out: torch.Tensor | None = None, group=None) -> torch.Tensor:
    """One collective per round: [world, gather_bytes] uint8, rank-major.  `local` is the rank's block prefix (for the GPU
    path a zero-copy view of the library's device block)."""
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world, local.numel()), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), local, group=group)
    return out


def unpack_block(block: np.ndarray, batch: int, kmax: int):
    """One rank's block (uint8) -> (infos, recs, top_id, top_p) numpy views."""
    lay = MetaLayout.make(batch, kmax)
    infos = block[: batch * 32].view(INFO_DTYPE)
    recs = block[lay.off_recs: lay.off_recs + batch * kmax * 32].view(REC_DTYPE)
    top_id = block[lay.off_top_id: lay.off_top_id + batch * kmax * 4].view(np.uint32)
    top_p = block[lay.off_top_p: lay.off_top_p + batch * kmax * 4].view(np.float32)
    return infos, recs, top_id, top_p


def unpack_round(gathered: torch.Tensor, round_idx: int, batch: int, kmax: int, with_identity: bool = False):
    """Host side: gathered [world, gather_bytes] -> {global frame index: (info, recs, truncated[, top_id, top_p])} in frame order.
    Frames whose blobs fall outside the gathered batch*kmax records are truncated and flagged.  top_id / top_p are per CROP; crop
    n is blob n of the rank's batch as long as no frame of the batch exceeds kmax blobs (tb_frame_info.status bit 2)."""
    g = gathered.cpu().numpy()
    world = g.shape[0]
    out = {}
    for rank in range(world):
        infos, recs, top_id, top_p = unpack_block(g[rank], batch, kmax)
        lo, _ = frame_range(round_idx, rank, world, batch)
        for i in range(batch):
            b0, n = int(infos[i]["blob_begin"]), int(infos[i]["n_blobs"])
            hi = min(b0 + n, len(recs))
            b0c = min(b0, hi)
            row = (infos[i], recs[b0c:hi], hi < b0 + n)
            if with_identity:
                row += (top_id[b0c:hi], top_p[b0c:hi])
            out[lo + i] = row
    return dict(sorted(out.items()))
