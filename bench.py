#!/usr/bin/env python
"""bench.py -- frames/s of the TRex hot path (bg-sub -> blobs -> crops -> VisualIdentification CNN) on
synthetic 1080p frames with 100 individuals (BASELINE.json metric / configs[2]).

    python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...   the reference's CPU algorithm (oracle port) on host cores

A step = one pass of the hot path over one batch of `--batch` synthetic frames per GPU.
  value  device-timed (CUDA events on the launching stream), inputs resident in HBM, max over ranks
  e2e    the same through the host-facing API: pinned host frames -> H2D -> kernels -> D2H of blob
         lists + identity probabilities, every step, wall clock bracketed by device syncs
Frames shard across ranks (weak scaling: every rank processes its own batch); the only collective is
one NCCL all-gather of the fixed-stride blob metadata per step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

H, W, N_INDIV, M_CLASSES = 1080, 1920, 100, 100
MAX_CROPS = 128
WORKLOAD = "synthetic 1920x1080 u8 gray, 100 individuals, bg-sub+threshold+CCL+80x80 crops+V118_3 CNN (random-init weights)"
# DRAM traffic per unit (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture divided by the
# units of that launch; profiles/r1_step_fp16_ncu_summary.txt and profiles/r1_step_bf16x3_ncu_summary.txt)
NCU_TRAFFIC = {"bf16x3": {"seg_rle": (267.87e6 + 6.47e6) / 128, "conv2": (507.7e6 + 385.1e6) / 4096, "conv3": (614.06e6 + 184.99e6) / 4096},
               "fp16": {"seg_rle": (267.87e6 + 6.47e6) / 128, "conv2": (253.88e6 + 175.66e6) / 4096, "conv3": (282.11e6 + 166.10e6) / 4096},
               "fp32": {"seg_rle": (267.87e6 + 6.47e6) / 128}}
MACS = {"conv1": 2.56e6, "conv2": 40.96e6, "conv3": 81.92e6, "fc1": 1.28e6, "head": 100.0 * M_CLASSES}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor=d.get("bf16_tflops_sustained", d["bf16_tflops"]), tensor_burst=d["bf16_tflops"],
                    src="measured (MEASURED_PEAKS.json; tensor = sustained bf16)")
    return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                     nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.005)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def bind_near_gpu(index):
    """Pin this rank's host threads (and so the first-touch placement of its pinned staging buffers) to the CPUs of
    the GPU's NUMA node; with 8 ranks on a 2-socket host a remote buffer halves the H2D rate.  Returns the CPU set."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        all_cpus = os.sched_getaffinity(0)
        nv.nvmlDeviceSetCpuAffinity(nv.nvmlDeviceGetHandleByIndex(index))
        near = os.sched_getaffinity(0)
        return all_cpus, near
    except Exception:  # noqa: BLE001
        return None, None


def make_inputs(n_frames, seed):
    from trex_b200.synthetic import BlobWorld
    world = BlobWorld(h=H, w=W, n_blobs=N_INDIV, seed=seed)
    return world.bg, world.frames(n_frames)


def weights():
    from trex_b200.weights import random_v118_3_state_dict
    return random_v118_3_state_dict(M_CLASSES, seed=0)


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's CPU algorithm on host cores
# --------------------------------------------------------------------------------------------------
def cpu_pipeline(bg, frames, sd, threads):
    """One pass of the reference's CPU path (restated in oracle/): segmentation + crops over all host
    threads (frames are independent), then V118_3 under torch CPU with the same thread count."""
    import torch
    from oracle import seg as oseg, vi as ovi
    torch.set_num_threads(threads)
    P = oseg.Params(detect_threshold=15, detect_size_filter=[(10.0, 100000.0)])
    nb, crops = oseg.segment_batch(frames, bg, P, crop_method=oseg.DIFF_ABSOLUTE, max_crops=MAX_CROPS, threads=threads)
    batch = np.concatenate([crops[f, :min(int(nb[f]), MAX_CROPS)] for f in range(len(frames))])[..., None]
    probs = ovi.predict(sd, batch)
    return int(nb.sum()), probs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import seg as oseg, vi as ovi
    oseg.build()
    threads = os.cpu_count() or 1
    sample = args.ref_frames
    bg, frames = make_inputs(sample, seed=1234)
    sd = ovi.scale_for_u8_inputs(ovi.init_state_dict(M_CLASSES, 1, 80, 80, seed=0))
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_pipeline(bg, frames, sd, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pipeline(bg, frames, sd, threads)
    dt = time.perf_counter() - t0
    fps = args.steps * sample / dt
    line = {
        "impl": "reference", "metric": "frames/sec (1080p, 100 indiv, bg-sub->blobs->CNN ID)", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 (seg) + f32 (CNN)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": sample, "note": "TRex cannot be built here (needs OpenCV C++/glaze): "
                   "this arm times the oracle port of its CPU algorithm (oracle/trex_oracle.c + torch CPU V118_3)"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} frames x {args.steps} steps, seg over {threads} pthreads + torch CPU CNN ({threads} threads)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import trex_b200

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    all_cpus, near_cpus = bind_near_gpu(local_rank) if not args.no_numa else (None, None)
    if world_size > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line (a nccl.conf may ask for the version banner)
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    pk = peaks()

    # ---- inputs: `pool` distinct batches resident in HBM (> L2 so no step is served from cache) ----
    n_src = min(B, 32)
    bg, src = make_inputs(n_src, seed=1234 + rank)
    CN, rgb8 = args.channels, args.encoding == "rgb8"
    if CN > 1:      # the colour variant of the same workload (not the headline): BGR(A) frames, cvtColor fused into K1
        from trex_b200.synthetic import to_color
        src = to_color(src, seed=rank, channels=CN)
        bg3 = to_color(bg, seed=99, channels=3)
        bg = bg3 if rgb8 else ((3735 * bg3[..., 0].astype(np.int64) + 19235 * bg3[..., 1].astype(np.int64) + 9798 * bg3[..., 2].astype(np.int64) + 16384) >> 15).astype(np.uint8)
    rng = np.random.default_rng(rank)
    pool = max(2, args.pool)
    host_batches = []
    for _ in range(pool):
        idx = rng.permutation(np.arange(B) % n_src)
        t = torch.empty((B, H, W) if CN == 1 else (B, H, W, CN), dtype=torch.uint8, pin_memory=True)
        t.numpy()[:] = src[idx]
        host_batches.append(t)
    dev_batches = [t.to(dev, non_blocking=True) for t in host_batches]
    torch.cuda.synchronize()

    settings = trex_b200.DetectSettings(meta_encoding=args.encoding)       # reference defaults: T=15, abs diff, size filter [10,100000)
    bs = trex_b200.BackgroundSubtraction(bg, settings=settings, max_batch=B, max_individuals=MAX_CROPS, device=local_rank, channels=CN)
    CI = 3 if rgb8 else 1
    if CI == 1:
        sd = weights()
    else:
        from trex_b200.weights import random_v118_3_state_dict
        sd = random_v118_3_state_dict(M_CLASSES, seed=0, channels=3)
    net = trex_b200.VINetwork(M_CLASSES, channels=CI, max_images=B * MAX_CROPS, device=local_rank, precision=args.precision)
    net.load_weights(sd)
    crops_p, ncrops_p, _, recs_p, infos_p = bs.device_results()
    probs = torch.empty((B * MAX_CROPS, M_CLASSES), dtype=torch.float32, device=dev)
    probs_host = torch.empty((B * MAX_CROPS, M_CLASSES), dtype=torch.float32, pin_memory=True)
    # fixed-stride metadata for the all-gather: the first B*MAX_CROPS blob records (32 B each) + headers
    from trex_b200 import sharding
    meta_bytes = sharding.meta_bytes(B, MAX_CROPS, with_identity=True)
    top_id = torch.zeros(B * MAX_CROPS, dtype=torch.int32, device=dev)
    top_p = torch.zeros(B * MAX_CROPS, dtype=torch.float32, device=dev)
    net.set_top1(top_id.data_ptr(), top_p.data_ptr())
    meta_all = torch.empty((world_size, meta_bytes), dtype=torch.uint8, device=dev) if world_size > 1 else None

    class _CudaBuf:      # zero-copy torch view of a device buffer owned by the C library
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}

    def as_tensor(ptr, nbytes):
        return torch.as_tensor(_CudaBuf(ptr, nbytes), device=dev)

    # one explicit (non-default) stream carries seg, CNN, copies and the collective; its handle goes to the C ABI
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)

    def step_device(i):
        fr = dev_batches[i % pool]
        bs.apply_device(fr.data_ptr(), B, stream.cuda_stream, fetch=False)
        net.predict_device(crops_p, B * MAX_CROPS, ncrops_p, probs.data_ptr(), 0, stream.cuda_stream)
        if world_size > 1:    # one collective per step: fixed-stride headers + blob records of every rank's frames
            sharding.all_gather_metadata(sharding.pack_metadata(as_tensor(infos_p, B * 32), as_tensor(recs_p, B * MAX_CROPS * 32), B, MAX_CROPS, top_id, top_p), out=meta_all)

    # e2e: two slots (seg handle + CNN handle + stream each) so the H2D copy of batch i+1 overlaps the
    # kernels of batch i; every step still moves its frames host->device and its results device->host.
    slots = []
    for k in range(2):
        st = torch.cuda.Stream(dev)
        if k == 0:
            sbs, snet = bs, net
        else:
            sbs = trex_b200.BackgroundSubtraction(bg, settings=settings, max_batch=B, max_individuals=MAX_CROPS, device=local_rank, channels=CN)
            snet = trex_b200.VINetwork(M_CLASSES, channels=CI, max_images=B * MAX_CROPS, device=local_rank, precision=args.precision)
            snet.load_weights(sd)
        s_id = torch.zeros(B * MAX_CROPS, dtype=torch.int32, device=dev)
        s_p = torch.zeros(B * MAX_CROPS, dtype=torch.float32, device=dev)
        slots.append(dict(bs=sbs, net=snet, stream=st, res=sbs.device_results(), pending=False, top_id=s_id, top_p=s_p,
                          probs=torch.empty((B * MAX_CROPS, M_CLASSES), dtype=torch.float32, device=dev),
                          probs_host=torch.empty((B * MAX_CROPS, M_CLASSES), dtype=torch.float32, pin_memory=True)))

    def e2e_submit(i):
        sl = slots[i % 2]
        sl["net"].set_top1(sl["top_id"].data_ptr(), sl["top_p"].data_ptr())
        sl["bs"].submit(host_batches[i % pool].numpy(), fetch=1)            # tb_seg_submit: H2D frames + kernels
        crops_q, ncrops_q = sl["res"][0], sl["res"][1]
        sl["net"].predict_device(crops_q, B * MAX_CROPS, ncrops_q, sl["probs"].data_ptr(), 0, sl["stream"].cuda_stream)
        with torch.cuda.stream(sl["stream"]):
            # identity probabilities back to the host (upper bound of rows: crops of this batch are not known yet)
            sl["probs_host"][:B * N_INDIV].copy_(sl["probs"][:B * N_INDIV], non_blocking=True)
            if world_size > 1:
                sharding.all_gather_metadata(sharding.pack_metadata(as_tensor(sl["res"][4], B * 32), as_tensor(sl["res"][3], B * MAX_CROPS * 32), B, MAX_CROPS, sl["top_id"], sl["top_p"]), out=meta_all)
        sl["pending"] = True

    def e2e_wait(i):
        sl = slots[i % 2]
        if not sl["pending"]:
            return 0, 0
        sl["bs"].wait()                                                     # tb_seg_wait: blob records, lines, pixels on the host
        sl["stream"].synchronize()
        sl["pending"] = False
        nb, nl, npx, nc = sl["bs"].totals()
        return B * H * W * CN, B * 32 + 16 + nb * 32 + nl * 8 + npx + B * N_INDIV * M_CLASSES * 4

    def barrier():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for i in range(max(args.warmup, 3)):
        step_device(i)
    barrier()
    bs.wait()
    tot = bs.totals()
    infos = as_tensor(infos_p, B * 32).cpu().numpy().view(np.uint32).reshape(B, 8)
    runs_per_batch = int(infos[:, 6].sum())

    # ---- timed region: value (device timed, inputs resident) ----
    bs.profile(True); net.profile(True)
    l0 = bs.launch_count() + net.launch_count()
    sampler = ClockSampler(local_rank); sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step_device(i)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = bs.launch_count() + net.launch_count() - l0
    seg_ms, seg_n = bs.kernel_ms()
    vi_ms, vi_n = net.kernel_ms()
    bs.profile(False); net.profile(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world_size * B * args.steps / (ms * 1e-3)

    # ---- e2e (host buffers in, host results out) ----
    torch.cuda.set_stream(torch.cuda.default_stream(dev))
    bs.wait()
    for sl in slots:
        sl["bs"].set_stream(sl["stream"].cuda_stream)
    for i in range(4):
        e2e_wait(i); e2e_submit(i)
    e2e_wait(0); e2e_wait(1)
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    for i in range(args.steps):
        a, b = e2e_wait(i)
        h2d += a; d2h += b
        e2e_submit(i)
    for i in range(2):
        a, b = e2e_wait(i)
        h2d += a; d2h += b
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world_size > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    e2e = world_size * B * args.steps / dt

    if rank == 0:
        n_crops = tot[3]
        per = {}
        for k, v in seg_ms.items():
            per[k] = v / max(seg_n, 1)
        for k, v in vi_ms.items():
            per[k] = v / max(args.steps, 1)          # vi events are per chunk; sum over the step
        total_k = sum(per.values())
        kern = {}
        alg_seg = B * W * H * CN + 8 * runs_per_batch    # frame read once + run records written
        kern["seg_rle"] = {"ms": per["seg_rle"], "share": per["seg_rle"] / total_k, "bound": "hbm",
                           "achieved": alg_seg / (per["seg_rle"] * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s"}
        for k in ("conv1", "conv2", "conv3", "fc1"):
            fl = 2 * MACS[k] * n_crops * (CI if k == "conv1" else 1)
            kern[k] = {"ms": per[k], "share": per[k] / total_k, "bound": "tensor",
                       "achieved": fl / (per[k] * 1e-3) / 1e12, "peak": pk["tensor"], "unit": "TFLOP/s"}
        for k in ("ccl_label", "blob_emit", "head"):
            kern[k] = {"ms": per[k], "share": per[k] / total_k}
        for v in kern.values():
            if "achieved" in v:
                v["frac"] = v["achieved"] / v["peak"]
        dom = max(("seg_rle", "conv1", "conv2", "conv3", "fc1"), key=lambda k: per[k])
        units = {"seg_rle": B, "conv2": n_crops, "conv3": n_crops}
        for k, per_unit in NCU_TRAFFIC[args.precision].items():
            if CN == 1 or k != "seg_rle":
                kern[k]["traffic"] = per_unit * units[k]          # bytes per step, from the committed ncu capture
        roof = {"kernel": dom, "bound": kern[dom]["bound"], "achieved": kern[dom]["achieved"], "peak": kern[dom]["peak"],
                "unit": kern[dom]["unit"], "frac": kern[dom]["frac"], "traffic": kern[dom].get("traffic"), "peak_source": pk["src"],
                "frac_of_burst_peak": (kern[dom]["achieved"] / pk["tensor_burst"]) if kern[dom]["bound"] == "tensor" else None,
                "note": "algorithmic FLOPs (2*MAC per crop)" + ("; the bf16x3 split issues 3 MMAs per k-step on top of that" if args.precision == "bf16x3" else "")}
        # cpu baseline on a bounded sample of the same workload (rank 0, N=1 only)
        cpu = None
        if world_size == 1 and not args.no_cpu and CN == 1:
            from oracle import seg as oseg, vi as ovi
            oseg.build()
            if all_cpus:
                os.sched_setaffinity(0, all_cpus)          # the CPU arm uses every host core again
            threads = os.cpu_count() or 1
            sd = ovi.scale_for_u8_inputs(ovi.init_state_dict(M_CLASSES, 1, 80, 80, seed=0))
            sample = src[:args.ref_frames]
            cpu_pipeline(bg, sample[:2], sd, threads)
            reps, t0 = 0, time.perf_counter()
            while reps < 3 or (time.perf_counter() - t0 < 10 and reps < 50):
                cpu_pipeline(bg, sample, sd, threads); reps += 1
            cdt = time.perf_counter() - t0
            cpu = {"value": reps * len(sample) / cdt, "unit": "frames/s", "cores": threads, "kind": "port",
                   "sample": f"{len(sample)} frames x {reps} reps of the same workload; oracle/trex_oracle.c over {threads} pthreads + torch CPU V118_3"}
        line = {
            "metric": "frames/sec (1080p, 100 indiv, bg-sub->blobs->CNN ID)", "value": value, "unit": "frames/s",
            "n_gpus": world_size, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 (seg) + " + {"bf16x3": "bf16x3 split, f32 accumulate (CNN)", "fp16": "f16 operands, f32 accumulate (CNN conv2/conv3; conv1, fc1 bf16x3)", "fp32": "f32 (CNN)"}[args.precision], "data": "synthetic",
            "config": {"workload": WORKLOAD if CN == 1 else WORKLOAD.replace("u8 gray", f"u8 x{CN} (BGR{'A' if CN == 4 else ''}), meta_encoding {args.encoding}"),
                       "frames_per_step_per_gpu": B, "crops_per_step_per_gpu": n_crops,
                       "blobs_per_step": tot[0], "l2": f"rotating pool of {pool} distinct batches ({pool * B * H * W * CN / 1e6:.0f} MB) > 126 MB L2",
                       "parallelism": f"frame-batch data parallel x{world_size}" + (", NCCL all-gather of blob metadata" if world_size > 1 else "")},
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps,
                    "timing": "wall clock bracketed by device syncs, max over ranks; 2 batches in flight (H2D of batch i+1 overlaps kernels of batch i)",
                    "host_cpus_near_gpu": len(near_cpus) if near_cpus else None},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernels": kern, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256, help="frames per step per GPU (256: K1 reaches 0.72 of the HBM peak, 0.64 at 128; the frame rate is the same)")
    ap.add_argument("--pool", type=int, default=4, help="distinct resident batches rotated through (defeats L2)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-frames", type=int, default=8, help="frames per step of the CPU arm / cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank to the CPUs of its GPU's NUMA node")
    ap.add_argument("--channels", type=int, default=1, choices=[1, 3, 4], help="bytes per pixel of the frames (3 BGR, 4 BGRA: colour variant, not the headline)")
    ap.add_argument("--encoding", default="gray", choices=["gray", "rgb8"], help="meta_encoding (rgb8 needs --channels 3|4; crops and conv1 then have 3 channels)")
    ap.add_argument("--precision", default="fp16", choices=["fp32", "bf16x3", "fp16"],
                    help="CNN arithmetic: fp32 CUDA cores, bf16x3 split (3 MMAs per k-step) or fp16 (1 MMA per k-step in conv2/conv3) on tcgen05")
    ap.add_argument("--individuals", type=int, default=100, help="blobs per frame = classes of the network (256: BASELINE config 4; not the headline)")
    ap.add_argument("--size", default="1920x1080", help="frame size WxH (3840x2160: BASELINE config 5's frames; not the headline)")
    args = ap.parse_args()
    global H, W, N_INDIV, M_CLASSES, MAX_CROPS, WORKLOAD
    if args.individuals != 100 or args.size != "1920x1080":
        W, H = (int(v) for v in args.size.split("x"))
        N_INDIV = M_CLASSES = args.individuals
        MAX_CROPS = (N_INDIV * 5 // 4 + 31) // 32 * 32
        MACS["head"] = 100.0 * M_CLASSES
        WORKLOAD = WORKLOAD.replace("1920x1080", f"{W}x{H}").replace("100 individuals", f"{N_INDIV} individuals")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
