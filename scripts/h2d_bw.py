#!/usr/bin/env python
"""Bare host->device copy bandwidth of this box: N processes (one per GPU of a subset), each copying a page-locked buffer of one
benchmark batch (256 x 1080p frames = 531 MB) to its GPU with cudaMemcpyAsync, all at the same time.  Prints one JSON line per
subset with GB/s per GPU and in aggregate, plus the PCIe / NUMA topology (nvidia-smi topo -m, NVML common ancestors), so that
bench.py's end-to-end scaling can be compared with the ceiling of the host's copy path ("what's weak" 13).

    python scripts/h2d_bw.py [--mb 531] [--reps 8] [--subsets "0;0,1;0,1,2,3;0,1,4,5;0,1,2,3,4,5,6,7"] [--pageable]
"""
import argparse
import json
import os
import subprocess
import sys
import time

import torch
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def worker(k, gpus, nbytes, reps, pageable, barrier, q):
    dev_index = gpus[k]
    torch.cuda.set_device(dev_index)
    try:
        import pynvml as nv
        nv.nvmlInit()
        nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(dev_index))
    except Exception:  # noqa: BLE001
        pass
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=not pageable)
    host.fill_(k + 1)
    dst = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{dev_index}")
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        dst.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    barrier.wait()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(reps):
            dst.copy_(host, non_blocking=True)
        e1.record(st)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    barrier.wait()
    q.put((dev_index, reps * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9, reps * nbytes / wall / 1e9))


def topology():
    out = {}
    try:
        out["nvidia_smi_topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
    except Exception as e:  # noqa: BLE001
        out["nvidia_smi_topo"] = f"unavailable: {e}"
    try:
        import pynvml as nv
        nv.nvmlInit()
        n = nv.nvmlDeviceGetCount()
        hs = [nv.nvmlDeviceGetHandleByIndex(i) for i in range(n)]
        out["common_ancestor"] = [[int(nv.nvmlDeviceGetTopologyCommonAncestor(hs[i], hs[j])) if i != j else 0 for j in range(n)] for i in range(n)]
        out["pci_bus"] = [nv.nvmlDeviceGetPciInfo(h).busId if isinstance(nv.nvmlDeviceGetPciInfo(h).busId, str) else nv.nvmlDeviceGetPciInfo(h).busId.decode() for h in hs]
        out["levels"] = "0 same board, 10 single PCIe switch, 20 multiple switches, 30 same host bridge, 40 same NUMA node, 50 across sockets"
    except Exception as e:  # noqa: BLE001
        out["nvml"] = f"unavailable: {e}"
    out["cpus"] = os.cpu_count()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=530.8)
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--subsets", default="")
    ap.add_argument("--pageable", action="store_true")
    args = ap.parse_args()
    n = torch.cuda.device_count()
    if args.subsets:
        subsets = [[int(v) for v in s.split(",")] for s in args.subsets.split(";") if s]
    else:
        subsets = [[0]]
        if n >= 2:
            subsets.append([0, 1])
        if n >= 4:
            subsets += [[0, 1, 2, 3]]
        if n >= 8:
            subsets += [[0, 1, 4, 5], [4, 5, 6, 7], [0, 2, 4, 6], list(range(8))]
    print(json.dumps({"topology": topology()}), flush=True)
    nbytes = int(args.mb * 1e6)
    ctx = mp.get_context("spawn")
    for gpus in subsets:
        if max(gpus) >= n:
            continue
        barrier, q = ctx.Barrier(len(gpus)), ctx.Queue()
        procs = [ctx.Process(target=worker, args=(k, gpus, nbytes, args.reps, args.pageable, barrier, q)) for k in range(len(gpus))]
        for p in procs:
            p.start()
        res = sorted(q.get(timeout=300) for _ in procs)
        for p in procs:
            p.join(timeout=60)
        print(json.dumps({"gpus": gpus, "source": "pageable" if args.pageable else "page-locked", "mb_per_copy": args.mb, "reps": args.reps,
                          "gbs_per_gpu": {str(g): round(a, 2) for g, a, _ in res}, "gbs_aggregate": round(sum(a for _, a, _ in res), 2),
                          "gbs_aggregate_wall": round(sum(w for _, _, w in res), 2)}), flush=True)


if __name__ == "__main__":
    main()
