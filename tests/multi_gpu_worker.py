"""Worker of tests/test_gpu_multi.py, launched with torchrun (one rank per GPU): every rank segments and identifies its own
frames, the metadata blocks the kernels wrote in place are all-gathered over NCCL, and EVERY rank checks that the gathered block
of every other rank equals what that rank holds (a second gather of the raw blocks through torch.distributed objects) and that
unpack_round returns all frames in order with the right records and identities."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trex_b200  # noqa: E402
from trex_b200 import sharding  # noqa: E402
from trex_b200.synthetic import BlobWorld  # noqa: E402
from trex_b200.weights import random_v118_3_state_dict  # noqa: E402


class _CudaBuf:
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, KMAX, M, ROUNDS = 6, 32, 24, 2
    world_gen = BlobWorld(h=272, w=480, n_blobs=20, seed=100 + rank, margin=30)
    bs = trex_b200.BackgroundSubtraction(world_gen.bg, settings=trex_b200.DetectSettings(), max_batch=B, max_individuals=KMAX, device=local)
    net = trex_b200.VINetwork(M, max_images=B * KMAX, device=local, precision="bf16x3")
    net.load_weights(random_v118_3_state_dict(M, seed=0))
    net.set_top1(*bs.top1_ptrs())
    meta = bs.metadata()
    lay = sharding.MetaLayout.from_c(meta)
    block = torch.as_tensor(_CudaBuf(meta.base, meta.gather_bytes), device=dev)
    crops_p, ncrops_p, _, _, _ = bs.device_results()
    probs = torch.zeros((B * KMAX, M), dtype=torch.float32, device=dev)
    stream, side = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    gathered = torch.empty((world, lay.gather_bytes), dtype=torch.uint8, device=dev)
    ok = True
    for rnd in range(ROUNDS):
        frames = world_gen.frames(B)
        fd = torch.from_numpy(frames).to(dev)
        bs.apply_device(fd.data_ptr(), B, stream.cuda_stream, fetch=1)
        net.predict_device(crops_p, B * KMAX, ncrops_p, probs.data_ptr(), 0, stream.cuda_stream)
        ev = torch.cuda.Event(); ev.record(stream)
        side.wait_event(ev)
        with torch.cuda.stream(side):                        # the collective runs off the compute stream, straight from the block
            sharding.all_gather_metadata(block, out=gathered)
        side.synchronize()
        bs.wait()
        local_block = block.cpu().numpy().copy()
        blocks = [None] * world
        dist.all_gather_object(blocks, local_block.tobytes())
        g = gathered.cpu()
        for r in range(world):
            ok &= g[r].numpy().tobytes() == blocks[r]
        out = sharding.unpack_round(g, rnd, B, KMAX, with_identity=True)
        ok &= list(out) == list(range(rnd * world * B, (rnd + 1) * world * B))
        # this rank's own frames: headers / records / identities against the host results of the C ABI and the probabilities
        lo = sharding.frame_range(rnd, rank, world, B)[0]
        p = probs.cpu().numpy()
        for i in range(B):
            info, recs, trunc, ids, ps = out[lo + i]
            _, hrecs, _, _ = bs.raw_result(i)
            ok &= (not trunc) and int(info["n_blobs"]) == len(hrecs) and np.array_equal(recs, hrecs)
            b0 = int(info["blob_begin"])
            ok &= np.array_equal(ids, p[b0:b0 + len(ids)].argmax(1)) and np.allclose(ps, p[b0:b0 + len(ids)].max(1), atol=1e-6)
            ok &= len(hrecs) > 0
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok))
    if rank == 0:
        print("MULTI_GPU_META_OK" if all(flags) else f"MULTI_GPU_META_FAIL {flags}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if all(flags) else 1)


if __name__ == "__main__":
    main()
