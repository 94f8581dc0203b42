"""Frame-batch data parallelism over the GPUs of one box.

Given a fixed background every frame is independent through segmentation, crops and the CNN
(BackgroundSubtraction::apply keeps no cross-frame state, T/python/BackgroundSubtraction.cpp:126-347),
so round k of a G-GPU job gives GPU g the frames  [k*G*B + g*B, k*G*B + (g+1)*B).  The only exchange is
one all-gather per round of the fixed-stride blob metadata (tb_frame_info[B] + the first B*Kmax
tb_blob_rec, 32 bytes each) so that the host tracker of every rank sees all frames in order.
Works with any torch.distributed backend (NCCL on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

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


def meta_bytes(batch: int, kmax: int, with_identity: bool = False) -> int:
    return batch * INFO_DTYPE.itemsize + batch * kmax * (REC_DTYPE.itemsize + (8 if with_identity else 0))


def pack_metadata(infos: torch.Tensor, recs: torch.Tensor, batch: int, kmax: int,
                  top_id: torch.Tensor | None = None, top_p: torch.Tensor | None = None) -> torch.Tensor:
    """Concatenate the per-frame headers, the first batch*kmax blob records and (optionally) the identity the CNN
    assigned to each of them (arg-max class uint32 + probability float32); uint8 tensors on any device."""
    parts = [infos.view(torch.uint8)[: batch * 32], recs.view(torch.uint8)[: batch * kmax * 32]]
    if top_id is not None:
        parts += [top_id.view(torch.uint8)[: batch * kmax * 4], top_p.view(torch.uint8)[: batch * kmax * 4]]
    return torch.cat(parts)


def all_gather_metadata(local: torch.Tensor, out: torch.Tensor | None = None, group=None) -> torch.Tensor:
    """One collective per round: [world, meta_bytes] uint8, rank-major."""
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world, local.numel()), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), local.contiguous(), group=group)
    return out


def unpack_round(gathered: torch.Tensor, round_idx: int, batch: int, kmax: int):
    """Host side: gathered [world, meta_bytes] -> {global frame index: (info, recs)} in frame order.
    Frames whose blob count exceeds the gathered records (more than kmax on average) are truncated and flagged."""
    g = gathered.cpu().numpy()
    world = g.shape[0]
    out = {}
    for rank in range(world):
        infos = g[rank, : batch * 32].view(INFO_DTYPE)
        recs = g[rank, batch * 32: batch * 32 + batch * kmax * 32].view(REC_DTYPE)
        lo, _ = frame_range(round_idx, rank, world, batch)
        for i in range(batch):
            b0, n = int(infos[i]["blob_begin"]), int(infos[i]["n_blobs"])
            hi = min(b0 + n, len(recs))
            out[lo + i] = (infos[i], recs[b0:hi], hi < b0 + n)
    return dict(sorted(out.items()))
