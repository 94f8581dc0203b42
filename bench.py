#!/usr/bin/env python
"""bench.py -- frames/s of the TRex hot path (bg-sub -> blobs -> crops -> VisualIdentification CNN) on synthetic frames.

    python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...   the reference's CPU algorithm (oracle port) on host cores
    python bench.py --config 4|5 ...                          BASELINE.json configs[3] / configs[4] (default 3 = configs[2], the metric's)

A step = one pass of the hot path over one batch of synthetic frames per GPU.
  value  device-timed (CUDA events on the launching stream), inputs resident in HBM, max over ranks
  e2e    the same through the host-facing C ABI: page-locked host frames (tb_host_alloc) -> tb_seg_submit (H2D inside) -> kernels ->
         D2H of blob lists + identity probabilities, every step, wall clock bracketed by device syncs; `e2e.pageable` is the same
         from ordinary (pageable) host memory
Both are reported for ALL tensor-core precisions of the CNN: the top-level value / e2e belong to --precision (default fp16c, the
library default: fp16 + e5m2 correction terms, within the 1e-3 logit tolerance on every weight set of tests/test_gpu_chain.py),
the others (bf16x3: the most accurate; fp16: the fastest, O(1) logits only) appear as value_<p> / e2e_<p>.
After the timed region one frame per rank is checked against the CPU oracle ("verified"); for N>1 the gathered metadata of every
rank is checked against what the rank produced ("meta_verified").
Frames shard across ranks (weak scaling: every rank processes its own batch); the only collective is one NCCL all-gather per step
of the metadata block the kernels write in place (no packing), on a side stream.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# BASELINE.json configs[2..4]: frame size, ellipses per frame, classes, crops gathered per frame, frames per step
CONFIGS = {
    3: dict(W=1920, H=1080, indiv=100, classes=100, kmax=128, batch=256, posture=False, name="configs[2]"),
    4: dict(W=1920, H=1080, indiv=256, classes=256, kmax=256, batch=128, posture=False, name="configs[3]"),
    5: dict(W=3840, H=2160, indiv=100, classes=100, kmax=128, batch=64, posture=True, name="configs[4]"),
}
MACS = {"conv1": 2.56e6, "conv2": 40.96e6, "conv3": 81.92e6, "fc1": 1.28e6}
# DRAM traffic per unit from `ncu --set full` captures (dram__bytes_read.sum + dram__bytes_write.sum of one launch / units of that
# launch): NOT measured in this run -- the capture each figure comes from is named next to it
NCU_TRAFFIC = {
    "fp16c": {"seg_rle": ((134.86e6 + 5.51e6) / 64, "profiles/r2_step_fp16c_head_ncu_summary.txt (B=64 capture, per frame)"),
              "conv2": ((507.65e6 + 378.80e6) / 4096, "profiles/r2_step_fp16c_head_ncu_summary.txt (4096-crop launch, per crop)"),
              "conv3": ((563.21e6 + 183.19e6) / 4096, "profiles/r2_step_fp16c_head_ncu_summary.txt (4096-crop launch, per crop)")},
    "bf16x3": {"seg_rle": ((267.87e6 + 6.47e6) / 128, "profiles/r1_step_bf16x3_ncu_summary.txt (B=128 capture, per frame)"),
               "conv2": ((507.7e6 + 385.1e6) / 4096, "profiles/r1_step_bf16x3_ncu_summary.txt (4096-crop launch, per crop)"),
               "conv3": ((614.06e6 + 184.99e6) / 4096, "profiles/r1_step_bf16x3_ncu_summary.txt (4096-crop launch, per crop)")},
    "fp16": {"seg_rle": ((267.87e6 + 6.47e6) / 128, "profiles/r1_step_fp16_ncu_summary.txt (B=128 capture, per frame)"),
             "conv2": ((253.88e6 + 175.66e6) / 4096, "profiles/r1_step_fp16_ncu_summary.txt (4096-crop launch, per crop)"),
             "conv3": ((282.11e6 + 166.10e6) / 4096, "profiles/r1_step_fp16_ncu_summary.txt (4096-crop launch, per crop)")},
}
DTYPE = {"fp16c": "u8 (seg) + f16 operands + e5m2 correction terms, f32 accumulate (CNN conv2/conv3; conv1, fc1 bf16x3)",
         "bf16x3": "u8 (seg) + bf16x3 split, f32 accumulate (CNN)",
         "fp16": "u8 (seg) + f16 operands, f32 accumulate (CNN conv2/conv3; conv1, fc1 bf16x3)",
         "fp32": "u8 (seg) + f32 (CNN)"}


def workload(cfg):
    w = (f"synthetic {cfg['W']}x{cfg['H']} u8 gray, {cfg['indiv']} moving ellipses per frame (overlapping ellipses merge: ~"
         f"{0.914 * cfg['indiv']:.0f} blobs per frame at 1080p), bg-sub+threshold+CCL+80x80 crops+V118_3 CNN (random-init weights, M={cfg['classes']})")
    return w + (" + posture (outline + midline per blob)" if cfg["posture"] else "")


def metric(cfg):
    return {3: "frames/sec (1080p, 100 indiv, bg-sub->blobs->CNN ID)", 4: "frames/sec (1080p, 256 indiv, bg-sub->blobs->CNN ID, frame-batch sharded)",
            5: "frames/sec (4K, 100 indiv, bg-sub->blobs->CNN ID + posture midline)"}[cfg["id"]]


def config_block(cfg):
    """Identical for our arm and the reference arm: what is computed, not how the run is sized."""
    return {"workload": workload(cfg), "baseline_config": cfg["name"], "frame_size": f"{cfg['W']}x{cfg['H']}", "individuals": cfg["indiv"],
            "classes": cfg["classes"], "detect_threshold": 15, "detect_size_filter": [10, 100000], "crop": "80x80x1 |bg-px|", "network": "v118_3"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), tensor_burst=d["bf16_tflops"],
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor_sustained=1400.0, tensor_burst=1400.0, src="fallback (B200_PROFILING.md)")


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, nvml_handle_fn):
        super().__init__(daemon=True)
        self.fn, self.samples, self.reasons, self.max_mhz, self._stop_evt = nvml_handle_fn, [], set(), None, threading.Event()

    def run(self):
        try:
            import pynvml as nv
            h = self.fn()
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


# --------------------------------------------------------------------------------------------------
# GPU selection: local rank -> CUDA device so that N < visible GPUs spreads over the PCIe host bridges
# --------------------------------------------------------------------------------------------------
def nvml_handle(cuda_index):
    """NVML handle of a CUDA device (NVML ignores CUDA_VISIBLE_DEVICES: match by PCI bus id)."""
    import pynvml as nv
    import torch
    nv.nvmlInit()
    p = torch.cuda.get_device_properties(cuda_index)
    try:
        bus = f"{p.pci_domain_id:08x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        return nv.nvmlDeviceGetHandleByPciBusId(bus.encode() if hasattr(bus, "encode") else bus)
    except Exception:  # noqa: BLE001
        return nv.nvmlDeviceGetHandleByIndex(cuda_index)


def device_order():
    """CUDA devices ordered so that consecutive local ranks sit under DIFFERENT PCIe host bridges / NUMA nodes (round robin over
    the groups nvmlDeviceGetTopologyCommonAncestor yields).  Returns (order, groups, why)."""
    import torch
    n = torch.cuda.device_count()
    try:
        import pynvml as nv
        hs = [nvml_handle(i) for i in range(n)]
        for level, name in ((nv.NVML_TOPOLOGY_HOSTBRIDGE, "pcie host bridge"), (nv.NVML_TOPOLOGY_NODE, "numa node")):
            groups = []
            for i in range(n):
                for g in groups:
                    if nv.nvmlDeviceGetTopologyCommonAncestor(hs[i], hs[g[0]]) <= level:
                        g.append(i)
                        break
                else:
                    groups.append([i])
            if 1 < len(groups) < n:
                order = [g[k] for k in range(max(len(g) for g in groups)) for g in groups if k < len(g)]
                return order, groups, name
        if n >= 4 and n % 2 == 0:
            # no PCIe structure visible (virtualised topology): HGX boards hang GPUs 0 .. n/2-1 and n/2 .. n-1 off different host
            # bridges -- profiles/r2_h2d_bw_8gpu.jsonl: {0,1,2,3} share 116 GB/s, {0,1,4,5} reach 210 GB/s -- so interleave the halves
            order = [g for k in range(n // 2) for g in (k, n // 2 + k)]
            return order, [list(range(n // 2)), list(range(n // 2, n))], "no PCIe structure reported: halves interleaved (h2d_bw table)"
        return list(range(n)), [list(range(n))], "one group (or all separate)"
    except Exception as e:  # noqa: BLE001
        return list(range(n)), [list(range(n))], f"nvml unavailable ({type(e).__name__})"


def bind_near_gpu(cuda_index):
    """Pin this rank's host threads (and so the first-touch placement of its staging buffers) to the CPUs near the GPU."""
    try:
        import pynvml as nv
        all_cpus = os.sched_getaffinity(0)
        nv.nvmlDeviceSetCpuAffinity(nvml_handle(cuda_index))
        return all_cpus, os.sched_getaffinity(0)
    except Exception:  # noqa: BLE001
        return None, None


def balanced_batches(batch, gbs, tolerance=0.9):
    """Frames per step of every rank in the end-to-end loop.  Equal shares (`batch`) unless the ranks' CONCURRENT host->device copy rates differ by
    more than 1 - tolerance: every step ends with one collective, so the rank with the slowest copy path would set the step time of all.  Then a
    rank's share follows its rate (the fastest keeps `batch`)."""
    top = max(gbs)
    if len(gbs) < 2 or top <= 0 or min(gbs) >= tolerance * top:
        return [batch] * len(gbs)
    return [max(1, min(batch, int(round(batch * g / top)))) for g in gbs]


def make_inputs(cfg, n_frames, seed):
    from trex_b200.synthetic import BlobWorld
    world = BlobWorld(h=cfg["H"], w=cfg["W"], n_blobs=cfg["indiv"], seed=seed)
    return world.bg, world.frames(n_frames)


def host_alloc(shape):
    """uint8 numpy array in page-locked memory from the library's own allocator (tb_host_alloc): what INTEGRATION.md's ImageMaker uses."""
    from trex_b200 import _capi
    n = int(np.prod(shape))
    p = C.c_void_p()
    _capi.check(_capi.lib().tb_host_alloc(n, C.byref(p)))
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (n,)).reshape(shape)


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port of the reference's CPU algorithm on host cores
# --------------------------------------------------------------------------------------------------
def cpu_pipeline(cfg, bg, frames, sd, threads):
    """One pass of the reference's CPU path (restated in oracle/): segmentation + crops over `threads` host threads (frames are
    independent), then V118_3 under torch CPU with the same thread count (+ the posture chain for config 5)."""
    import torch
    from oracle import seg as oseg, vi as ovi
    torch.set_num_threads(threads)
    P = oseg.Params(detect_threshold=15, detect_size_filter=[(10.0, 100000.0)])
    nb, crops = oseg.segment_batch(frames, bg, P, crop_method=oseg.DIFF_ABSOLUTE, max_crops=cfg["kmax"], threads=threads)
    batch = np.concatenate([crops[f, :min(int(nb[f]), cfg["kmax"])] for f in range(len(frames))])[..., None]
    probs = ovi.predict(sd, batch)
    if cfg["posture"]:
        from oracle import posture as opost
        for f in range(len(frames)):            # posture of every blob: longest outline -> resample -> midline (single thread, as the tracker's per-individual work)
            blobs = oseg.segment_frame(frames[f], bg, P)
            for k in range(len(blobs)):
                opost.calculate_midline(oseg.outline_resample(oseg.longest_outline(blobs.blob(k)[0]), 1.0))
    return int(nb.sum()), probs


def cpu_weights(cfg):
    from oracle import vi as ovi
    return ovi.scale_for_u8_inputs(ovi.init_state_dict(cfg["classes"], 1, 80, 80, seed=0))


def time_cpu(cfg, bg, frames, sd, threads, budget_s, min_reps=2, max_reps=50):
    cpu_pipeline(cfg, bg, frames[:1], sd, threads)
    reps, t0 = 0, time.perf_counter()
    while reps < min_reps or (time.perf_counter() - t0 < budget_s and reps < max_reps):
        cpu_pipeline(cfg, bg, frames, sd, threads)
        reps += 1
    return reps * len(frames) / (time.perf_counter() - t0), reps


def cv2_cross_check(cfg, bg, frames, budget_s=3.0):
    """The segmentation stage alone through OpenCV (what the reference itself calls: absdiff, threshold, bitwise_and, then 8-connected
    labelling -- here cv2.connectedComponentsWithStats instead of TRex's run-based CPULabeling): a plausibility check of the port's speed."""
    try:
        import cv2
    except Exception:  # noqa: BLE001
        return None
    cv2.setNumThreads(1)
    reps, t0, nb = 0, time.perf_counter(), 0
    while reps < 1 or time.perf_counter() - t0 < budget_s:
        for f in frames:
            d = cv2.absdiff(f, bg)
            _, m = cv2.threshold(d, 15, 255, cv2.THRESH_BINARY)
            out = cv2.bitwise_and(m, f)
            n, _, stats, _ = cv2.connectedComponentsWithStats((out > 0).astype(np.uint8), connectivity=8)
            nb += int(((stats[1:, cv2.CC_STAT_AREA] >= 10) & (stats[1:, cv2.CC_STAT_AREA] < 100000)).sum())
        reps += 1
    return {"seg_only_fps_1_thread": reps * len(frames) / (time.perf_counter() - t0), "blobs_per_frame": nb / (reps * len(frames)),
            "what": "cv2 absdiff+threshold+bitwise_and+connectedComponentsWithStats(8), 1 thread; segmentation stage only"}


def compiled_reference_check(cfg, bg, frames, budget_s=3.0):
    """The segmentation stage through the REFERENCE'S OWN code (oracle/ref_detect_leg.py: BackgroundSubtraction.cpp + RawProcessing.cpp + CPULabeling compiled
    unmodified, real OpenCV behind a callback bridge), one thread, in a CHILD PROCESS so that nothing in that library can cost the bench line.  A reported
    baseline next to the port's, never the thing measured."""
    import subprocess
    import tempfile
    try:
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "sample.npz")
            np.savez(path, bg=bg, frames=np.asarray(frames))
            r = subprocess.run([sys.executable, "-m", "oracle.ref_detect_leg", path, str(budget_s)], cwd=ROOT, capture_output=True, text=True, timeout=120)
        if r.returncode != 0 or not r.stdout.strip():
            return None
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:  # noqa: BLE001 -- an optional leg must never cost the bench line
        return None


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import seg as oseg
    oseg.build()
    threads = os.cpu_count() or 1
    sample = args.ref_frames
    bg, frames = make_inputs(cfg, sample, seed=1234)
    sd = cpu_weights(cfg)
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_pipeline(cfg, bg, frames, sd, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pipeline(cfg, bg, frames, sd, threads)
    dt = time.perf_counter() - t0
    fps = args.steps * sample / dt
    line = {
        "impl": "reference", "metric": metric(cfg), "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 (seg) + f32 (CNN)", "data": "synthetic",
        "config": config_block(cfg),
        "run": {"frames_per_step": sample, "note": "TRex as a whole cannot be built here (needs OpenCV C++ / glaze): this arm times the oracle port of its CPU algorithm (oracle/trex_oracle.c + "
                "torch CPU V118_3; the CNN is nine tenths of the time) on a bounded sample of the same workload; cpu_baseline.compiled_reference times the segmentation stage "
                "through the reference's own BackgroundSubtraction.cpp / RawProcessing.cpp / CPULabeling, compiled unmodified, with the real OpenCV"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "cpu": cpu_model(),
                         "sample": f"{sample} frames x {args.steps} steps, seg over {threads} pthreads + torch CPU CNN ({threads} threads)",
                         # the segmentation stage through the reference's OWN compiled BackgroundSubtraction::apply (oracle/_ref/libref_detect.so), for comparison with the port's
                         "compiled_reference": compiled_reference_check(cfg, bg, frames[:4])},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
class _CudaBuf:      # zero-copy torch view of a device buffer owned by the C library
    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist
    import trex_b200
    from trex_b200 import sharding
    from trex_b200.weights import random_v118_3_state_dict

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU baseline)")
    order, groups, order_why = device_order() if not args.no_topo else (list(range(torch.cuda.device_count())), [], "disabled")
    dev_index = order[local_rank % len(order)]
    torch.cuda.set_device(dev_index)
    dev = torch.device("cuda", dev_index)
    all_cpus, near_cpus = bind_near_gpu(dev_index) if not args.no_numa else (None, None)
    if world_size > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line (a nccl.conf may ask for the version banner)
        dist.init_process_group("nccl", device_id=dev)
    H, W, B, KMAX, M = cfg["H"], cfg["W"], args.batch or cfg["batch"], cfg["kmax"], cfg["classes"]
    pk = peaks()
    precisions = [args.precision] + [p for p in ("fp16c", "bf16x3", "fp16") if p != args.precision and not args.single_precision]

    def as_tensor(ptr, nbytes):
        return torch.as_tensor(_CudaBuf(ptr, nbytes), device=dev)

    # ---- inputs: `pool` distinct batches resident in HBM (> L2 so no step is served from cache) ----
    n_src = min(B, 32)
    bg, src = make_inputs(cfg, n_src, seed=1234 + rank)
    rng = np.random.default_rng(rank)
    pool = max(2, args.pool)
    host_batches, host_idx = [], []
    for _ in range(pool):
        idx = rng.permutation(np.arange(B) % n_src)
        t = host_alloc((B, H, W))                      # page-locked through the C ABI (tb_host_alloc)
        t[:] = src[idx]
        host_batches.append(t); host_idx.append(idx)
    dev_batches = [torch.from_numpy(t).to(dev, non_blocking=True) for t in host_batches]
    torch.cuda.synchronize()

    settings = trex_b200.DetectSettings()       # reference defaults: T=15, abs diff, size filter [10,100000)
    sd = random_v118_3_state_dict(M, seed=0)
    main = torch.cuda.Stream(dev)
    side = torch.cuda.Stream(dev)               # carries the all-gather: off the compute stream
    torch.cuda.set_stream(main)

    class Slot:
        pass

    slots = []
    for k in range(max(2, args.slots)):
        s = Slot()
        s.bs = trex_b200.BackgroundSubtraction(bg, settings=settings, max_batch=B, max_individuals=KMAX, device=dev_index)
        s.res = s.bs.device_results()
        s.meta = s.bs.metadata()
        s.lay = sharding.MetaLayout.from_c(s.meta)
        s.block = as_tensor(s.meta.base, s.meta.gather_bytes)            # zero-copy view of the block the kernels write
        s.top = s.bs.top1_ptrs()
        s.nets = {}
        s.probs = torch.empty((B * KMAX, M), dtype=torch.float32, device=dev)
        s.logits = torch.empty((B * KMAX, M), dtype=torch.float32, device=dev)
        s.probs_host = torch.empty((B * KMAX, M), dtype=torch.float32, pin_memory=True)
        s.stream = torch.cuda.Stream(dev)
        s.meta_all = torch.empty((world_size, s.meta.gather_bytes), dtype=torch.uint8, device=dev) if world_size > 1 else None
        s.ev_done, s.ev_gathered = torch.cuda.Event(), torch.cuda.Event()
        s.pending = s.used = False
        slots.append(s)

    def net_of(s, precision):
        if precision not in s.nets:
            n = trex_b200.VINetwork(M, max_images=B * KMAX, device=dev_index, precision=precision)
            n.load_weights(sd)
            n.set_top1(*s.top)
            s.nets[precision] = n
        return s.nets[precision]

    def gather(s, stream):
        """The one collective of a step, on the side stream, straight from the block the kernels wrote."""
        s.ev_done.record(stream)
        side.wait_event(s.ev_done)
        with torch.cuda.stream(side):
            sharding.all_gather_metadata(s.block, out=s.meta_all)
            s.ev_gathered.record(side)

    def step_device(i, precision, logits=False):
        s = slots[i % 2]
        s.used = True
        if world_size > 1:
            main.wait_event(s.ev_gathered)          # the block of this slot was gathered two steps ago
        s.bs.apply_device(dev_batches[i % pool].data_ptr(), B, main.cuda_stream, fetch=0)
        net_of(s, precision).predict_device(s.res[0], B * KMAX, s.res[1], s.probs.data_ptr(), s.logits.data_ptr() if logits else 0, main.cuda_stream)
        if cfg["posture"]:
            s.bs.posture_async(1.0, normalize=True, fetch=0)      # outlines -> midlines -> normalised midlines, behind the batch on `main`
        if world_size > 1:
            gather(s, main)

    def barrier():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def measure_device(precision):
        for i in range(max(args.warmup, 3)):
            step_device(i, precision)
        barrier()
        slots[0].bs.wait()
        tot = slots[0].bs.totals()
        infos = as_tensor(slots[0].res[4], B * 32).cpu().numpy().view(np.uint32).reshape(B, 8)
        runs_per_batch = int(infos[:, 6].sum())
        for s in slots:
            s.bs.profile(True); net_of(s, precision).profile(True)
        l0 = sum(s.bs.launch_count() + net_of(s, precision).launch_count() for s in slots)
        sampler = ClockSampler(lambda: nvml_handle(dev_index)); sampler.start()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for i in range(args.steps):
            step_device(i, precision)
        if world_size > 1:
            for s in slots:
                main.wait_event(s.ev_gathered)
        e1.record(main)
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop()
        launches = sum(s.bs.launch_count() + net_of(s, precision).launch_count() for s in slots) - l0
        per = {}
        n_seg = 0
        for s in slots:
            a, n = s.bs.kernel_ms(); n_seg += n
            b, _ = net_of(s, precision).kernel_ms()
            for k, v in {**a, **b}.items():
                per[k] = per.get(k, 0.0) + v
            s.bs.profile(False); net_of(s, precision).profile(False)
        if cfg["posture"]:
            for s in slots:
                for k, v in s.bs.posture_ms()[0].items():
                    per[k] = per.get(k, 0.0) + v
        per = {k: v / max(n_seg, 1) for k, v in per.items()}          # ms per step (vi events are per chunk, summed over the step)
        ms = max_over_ranks(ms)
        return dict(ms=ms, value=world_size * B * args.steps / (ms * 1e-3), clocks=clocks, launches=int(launches), per=per,
                    tot=tot, runs_per_batch=runs_per_batch)

    # ---- e2e: --slots slots (seg handle + CNN handle + stream each) so the H2D copy of batch i+1 overlaps the kernels of batch i; with
    # three, the copy engine already holds the next batch while the host fetches the results of the previous one.
    # Every step still moves its frames host->device and its results device->host ----
    NS = len(slots)

    def e2e_submit(i, precision, batches):
        s = slots[i % NS]
        s.used = True
        s.bs.submit(batches[i % len(batches)][:B_e], fetch=1)               # tb_seg_submit: H2D frames + kernels
        net_of(s, precision).predict_device(s.res[0], B * KMAX, s.res[1], s.probs.data_ptr(), 0, s.stream.cuda_stream)
        if cfg["posture"]:
            s.bs.posture_async(1.0, normalize=True, fetch=1)      # results: midline records + normalised midlines to the host
        with torch.cuda.stream(s.stream):
            # identity probabilities back to the host (upper bound of rows: the crop count of this batch is not known yet)
            s.probs_host[:B_e * cfg["indiv"]].copy_(s.probs[:B_e * cfg["indiv"]], non_blocking=True)
        if world_size > 1:
            gather(s, s.stream)
        s.pending = True

    def e2e_wait(i):
        s = slots[i % NS]
        if not s.pending:
            return 0, 0
        s.bs.wait()                                                          # tb_seg_wait: blob records, lines, pixels on the host
        if cfg["posture"]:
            s.bs.posture_wait()                                              # tb_seg_posture_wait: midlines on the host
        s.stream.synchronize()
        if world_size > 1:
            s.ev_gathered.synchronize()
        s.pending = False
        nb, nl, npx, nc = s.bs.totals()
        return B_e * H * W, B_e * 32 + 16 + nb * 32 + nl * 8 + npx + B_e * cfg["indiv"] * M * 4 + (nb * (16 + 32 + 25 * 16) if cfg["posture"] else 0)

    def measure_e2e(precision, batches, steps):
        for s in slots:
            if s.used:
                s.bs.wait()
            s.bs.set_stream(s.stream.cuda_stream)
        for i in range(2 * NS):
            e2e_wait(i); e2e_submit(i, precision, batches)
        for i in range(NS):
            e2e_wait(i)
        barrier()
        t0 = time.perf_counter()
        h2d = d2h = 0
        for i in range(steps):
            a, b = e2e_wait(i); h2d += a; d2h += b
            e2e_submit(i, precision, batches)
        for i in range(steps, steps + NS):
            a, b = e2e_wait(i); h2d += a; d2h += b
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        for s in slots:
            s.bs.set_stream(0)
        return dict(value=sum(B_e_all) * steps / dt, h2d=h2d // steps, d2h=d2h // steps, s_per_step=dt / steps)

    def bare_h2d():
        """GB/s of this rank's H2D copies alone (page-locked source), all ranks copying at the same time."""
        dst = dev_batches[0]
        srcs = [torch.from_numpy(t) for t in host_batches]
        with torch.cuda.stream(main):
            dst.copy_(srcs[0], non_blocking=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 6
        e0.record(main)
        with torch.cuda.stream(main):
            for r in range(reps):
                dst.copy_(srcs[r % len(srcs)], non_blocking=True)
        e1.record(main)
        barrier()
        gbs = reps * dst.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9
        with torch.cuda.stream(main):
            dst.copy_(srcs[0], non_blocking=True)          # restore batch 0
        torch.cuda.synchronize()
        if world_size > 1:
            out = [None] * world_size
            dist.all_gather_object(out, (dev_index, round(gbs, 2)))
            return out
        return [(dev_index, round(gbs, 2))]

    def verify(precision):
        """One frame of this rank against the CPU oracle: blob list, crops, logits / probabilities of its crops (abs 1e-3)."""
        from oracle import seg as oseg, vi as ovi
        s = slots[0]
        s.bs.apply_device(dev_batches[0].data_ptr(), B, main.cuda_stream, fetch=2)
        net_of(s, precision).predict_device(s.res[0], B * KMAX, s.res[1], s.probs.data_ptr(), s.logits.data_ptr(), main.cuda_stream)
        s.bs.wait(); main.synchronize()
        f = (7 * rank + 3) % B
        frame = src[host_idx[0][f]]
        P = oseg.Params(detect_threshold=15, detect_size_filter=[(10.0, 100000.0)])
        ref = oseg.segment_frame(frame, bg, P)
        got = [(b.lines.tobytes(), b.pixels.tobytes()) for b in s.bs.result(f)]
        ok = got == ref.as_list()
        crops, _ = s.bs.crops()
        c0 = sum(min(s.bs.frame_info(j).n_blobs, KMAX) for j in range(f))
        n = min(len(ref), KMAX)
        exp = np.stack([oseg.crop_blob(*ref.blob(k), bg, oseg.DIFF_ABSOLUTE) for k in range(n)]) if n else np.zeros((0, 80, 80), np.uint8)
        ok = ok and np.array_equal(crops[c0:c0 + n], exp)
        sdo = {k: v for k, v in sd.items()}
        dl = float(np.abs(s.logits[c0:c0 + n].cpu().numpy() - ovi.forward_logits(sdo, exp[..., None])).max()) if n else 0.0
        dp = float(np.abs(s.probs[c0:c0 + n].cpu().numpy() - ovi.predict(sdo, exp[..., None])).max()) if n else 0.0
        ok = ok and dl < 1e-3 and dp < 1e-3
        out = dict(ok=bool(ok), frame=int(f), blobs=len(ref), max_dlogit=dl, max_dprob=dp)
        if cfg["posture"]:                      # the posture chain of that frame's blobs against the oracle: raw midline bit-exact, normalised midline to 1e-4
            from oracle import posture as opost
            s.bs.posture_async(1.0, normalize=True, fetch=2)
            s.bs.posture_wait()
            pr = s.bs.posture_result()
            b0 = s.bs.frame_info(f).blob_begin
            okp, n_mid = True, 0
            for k in range(len(ref)):
                so, ns, tail, head = (int(v) for v in pr["midlines"][b0 + k])
                nr = pr["normalized"][b0 + k]
                try:
                    rs, rt, rh, _ = opost.calculate_midline(oseg.outline_resample(oseg.longest_outline(ref.blob(k)[0]), 1.0))
                except ValueError:
                    okp = okp and ns == 0
                    continue
                okp = okp and (tail, head) == (rt, rh) and np.array_equal(pr["segments"][so:so + ns], rs)
                pp, _, _, _ = opost.post_process(rs, tail=rt, head=rh)
                nm = opost.normalize(pp)
                if nm is None:
                    okp = okp and int(nr["n_points"]) == 0
                else:
                    okp = okp and int(nr["n_points"]) == 25 and float(np.abs(pr["norm_points"][b0 + k] - nm[0]).max()) < 1e-4 and abs(float(nr["len"]) - nm[1]) < 1e-3
                    n_mid += 1
            out["posture_ok"] = bool(okp); out["midlines_checked"] = n_mid
            out["ok"] = bool(ok and okp)
        return out

    def verify_meta():
        """N>1: every rank's gathered block equals what the rank produced (checksums through a second, tiny gather) and unpacks into
        all frames in order with rank 0's own headers / records / identities."""
        s = slots[0]
        main.synchronize()
        gather(s, main)
        s.ev_gathered.synchronize()
        local = s.block.cpu().numpy()
        mine = hashlib.sha1(local.tobytes()).hexdigest()
        sums = [None] * world_size
        dist.all_gather_object(sums, mine)
        g = s.meta_all.cpu()
        ok = all(hashlib.sha1(g[r].numpy().tobytes()).hexdigest() == sums[r] for r in range(world_size))
        frames = sharding.unpack_round(g, 0, B, KMAX, with_identity=True)
        ok = ok and list(frames) == list(range(world_size * B))
        infos, recs, top_id, top_p = sharding.unpack_block(local, B, KMAX)
        lo = sharding.frame_range(0, rank, world_size, B)[0]
        for i in (0, B // 2, B - 1):
            info, r, trunc, ids, ps = frames[lo + i]
            b0, n = int(infos[i]["blob_begin"]), int(infos[i]["n_blobs"])
            ok = ok and int(info["n_blobs"]) == n and np.array_equal(r, recs[b0:b0 + len(r)]) and np.array_equal(ids, top_id[b0:b0 + len(ids)])
            ok = ok and n > 0 and bool((ps > 0).all()) and not trunc
        flags = [None] * world_size
        dist.all_gather_object(flags, bool(ok))
        return all(flags)

    # ------------------------------------------------------------------------------------------
    res, verified = {}, {}
    for p in precisions:
        res[p] = measure_device(p)
        verified[p] = verify(p)
    meta_ok = verify_meta() if world_size > 1 else None
    h2d_gbs = bare_h2d()
    # frames per step of every rank in the end-to-end loop (list ordered by rank): equal unless the host's copy paths are not
    B_e_all = [B] * world_size if args.no_balance else balanced_batches(B, [g for _, g in h2d_gbs])
    B_e = B_e_all[rank]
    torch.cuda.set_stream(torch.cuda.default_stream(dev))
    e2e = {p: measure_e2e(p, host_batches, args.steps) for p in precisions}
    pageable = None
    if not args.no_pageable:
        pg = [np.array(host_batches[k]) for k in range(2)]                  # ordinary (pageable) copies of two batches
        pageable = measure_e2e(precisions[0], pg, max(3, args.steps // 4))
    all_ok = [None] * world_size
    if world_size > 1:
        dist.all_gather_object(all_ok, {p: v["ok"] for p, v in verified.items()})
    else:
        all_ok = [{p: v["ok"] for p, v in verified.items()}]

    if rank == 0:
        p0 = precisions[0]
        r0 = res[p0]
        n_crops, per = r0["tot"][3], r0["per"]
        region_s = r0["ms"] * 1e-3
        tensor_peak = pk["tensor_burst"] if region_s < 1.0 else pk["tensor_sustained"]

        def kernel_table(r):
            per, total_k, kern = r["per"], sum(r["per"].values()), {}
            alg_seg = B * W * H + 8 * r["runs_per_batch"]    # frame read once + run records written
            kern["seg_rle"] = {"ms": per["seg_rle"], "share": per["seg_rle"] / total_k, "bound": "hbm",
                               "achieved": alg_seg / (per["seg_rle"] * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s"}
            for k in ("conv1", "conv2", "conv3", "fc1"):
                fl = 2 * MACS[k] * r["tot"][3]
                kern[k] = {"ms": per[k], "share": per[k] / total_k, "bound": "tensor", "achieved": fl / (per[k] * 1e-3) / 1e12,
                           "peak": tensor_peak, "unit": "TFLOP/s"}
            for k in ("ccl_label", "blob_emit", "head") + (("outlines", "midlines") if cfg["posture"] else ()):
                kern[k] = {"ms": per[k], "share": per[k] / total_k}
            for v in kern.values():
                if "achieved" in v:
                    v["frac"] = v["achieved"] / v["peak"]
            # tensor-pipe slots per algorithmic MAC: conv2 / conv3 follow the precision; conv1 multiplies exact u8 pixels with hi / lo weights (2), fc1 is bf16x3 (3)
            slots = {"conv1": 2, "conv2": {"fp16": 1, "fp16c": 2, "bf16x3": 3}[p0], "conv3": {"fp16": 1, "fp16c": 2, "bf16x3": 3}[p0], "fc1": 3}
            for k, n in slots.items():
                kern[k]["mma_slots_per_mac"] = n
                kern[k]["frac_in_issued_slots"] = kern[k]["frac"] * n
            cnn_ms = sum(per[k] for k in ("conv1", "conv2", "conv3", "fc1", "head"))
            cnn_fl = r["tot"][3] * (2 * sum(MACS.values()) + 2 * 100 * M)
            return kern, {"tflops_algorithmic": cnn_fl / (cnn_ms * 1e-3) / 1e12, "frac_of_burst_peak": cnn_fl / (cnn_ms * 1e-3) / 1e12 / pk["tensor_burst"],
                          "ms": cnn_ms}

        kern, cnn = kernel_table(r0)
        if cfg["posture"]:
            dom = "seg_rle"       # config 5's headline roofline is the segmentation kernel against the HBM peak (BASELINE configs[4])
        else:
            dom = max(("seg_rle", "conv1", "conv2", "conv3", "fc1"), key=lambda k: per[k])
        units = {"seg_rle": B, "conv2": n_crops, "conv3": n_crops}
        traffic, traffic_src = None, None
        if dom in NCU_TRAFFIC.get(p0, {}) and cfg["id"] == 3:
            traffic = NCU_TRAFFIC[p0][dom][0] * units[dom]
            traffic_src = "ncu capture, scaled to this launch: " + NCU_TRAFFIC[p0][dom][1]
        roof = {"kernel": dom, "bound": kern[dom]["bound"], "achieved": kern[dom]["achieved"], "peak": kern[dom]["peak"], "unit": kern[dom]["unit"],
                "frac": kern[dom]["frac"], "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": pk["src"] + (f"; tensor peak = {'burst' if region_s < 1.0 else 'sustained'} bf16 figure (timed region {region_s:.2f} s)" if kern[dom]["bound"] == "tensor" else ""),
                "frac_of_burst_peak": kern[dom]["achieved"] / pk["tensor_burst"] if kern[dom]["bound"] == "tensor" else None,
                "frac_of_sustained_peak": kern[dom]["achieved"] / pk["tensor_sustained"] if kern[dom]["bound"] == "tensor" else None,
                "cnn_whole": cnn,
                # tensor-pipe slots the kernel occupies per algorithmic MAC: 1 (fp16), 2 (fp16c: fp16 K=16 + e5m2 K=32), 3 (bf16x3)
                "mma_slots_per_mac": {"fp16": 1, "fp16c": 2, "bf16x3": 3}[p0] if kern[dom]["bound"] == "tensor" else None,
                "frac_of_peak_in_issued_slots": (kern[dom]["frac"] * {"fp16": 1, "fp16c": 2, "bf16x3": 3}[p0]) if kern[dom]["bound"] == "tensor" else None,
                "note": "algorithmic FLOPs (2*MAC per crop)" + ("; the bf16x3 split issues 3 MMAs per k-step on top of that" if p0 == "bf16x3" else
                                                                 "; fp16c issues 2 MMA slots per k-step (fp16 K=16 + e5m2 K=32) on top of that" if p0 == "fp16c" else "")}
        # cpu baseline on a bounded sample of the same workload (rank 0, N=1 only): all cores, one thread, and an OpenCV cross-check
        cpu = None
        if world_size == 1 and not args.no_cpu:
            from oracle import seg as oseg
            oseg.build()
            if all_cpus:
                os.sched_setaffinity(0, all_cpus)          # the CPU arm uses every host core again
            threads = os.cpu_count() or 1
            sdc = cpu_weights(cfg)
            sample = src[:args.ref_frames]
            fps_all, reps_all = time_cpu(cfg, bg, sample, sdc, threads, 10.0)
            fps_1, reps_1 = time_cpu(cfg, bg, sample[:max(2, args.ref_frames // 4)], sdc, 1, 8.0, min_reps=1)
            cpu = {"value": fps_all, "unit": "frames/s", "cores": threads, "kind": "port", "cpu": cpu_model(),
                   "sample": f"{len(sample)} frames x {reps_all} reps of the same workload; oracle/trex_oracle.c over {threads} pthreads + torch CPU V118_3",
                   "one_thread": {"value": fps_1, "cores": 1, "sample": f"{max(2, args.ref_frames // 4)} frames x {reps_1} reps"},
                   "opencv_cross_check": cv2_cross_check(cfg, bg, sample[:4]),
                   "compiled_reference": compiled_reference_check(cfg, bg, sample[:4])}
        line = {
            "metric": metric(cfg), "value": r0["value"], "unit": "frames/s",
            "n_gpus": world_size, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": r0["ms"] / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE[p0], "precision": p0, "data": "synthetic",
            "config": config_block(cfg),
            "run": {"frames_per_step_per_gpu": B, "crops_per_step_per_gpu": n_crops, "blobs_per_step": r0["tot"][0], "blobs_per_frame": r0["tot"][0] / B,
                    "l2": f"rotating pool of {pool} distinct batches ({pool * B * H * W / 1e6:.0f} MB) > 126 MB L2",
                    "parallelism": f"frame-batch data parallel x{world_size}" + (", one NCCL all-gather of the in-place metadata block per step on a side stream" if world_size > 1 else ""),
                    "gpu_order": {"cuda_devices": order[:world_size], "groups": groups, "by": order_why}},
            "e2e": {"value": e2e[p0]["value"], "unit": "frames/s", "h2d_bytes_per_step": e2e[p0]["h2d"], "d2h_bytes_per_step": e2e[p0]["d2h"],
                    "timing": f"wall clock bracketed by device syncs, max over ranks, pipeline fill and drain inside; {max(2, args.slots)} batches in flight (H2D of batch i+1 overlaps kernels of batch i)",
                    "source": "page-locked host buffers from tb_host_alloc",
                    "pageable": ({"value": pageable["value"], "unit": "frames/s", "source": "ordinary pageable host memory (numpy)"} if pageable else None),
                    "frames_per_step_per_rank": B_e_all,
                    "sharding": ("frames per rank follow the ranks' concurrent H2D rates (h2d_gbs_per_rank differ by more than 10 %: one collective per step "
                                 "would make the slowest copy path set every rank's step time); h2d/d2h_bytes_per_step are rank 0's"
                                 if len(set(B_e_all)) > 1 else "equal frames per rank"),
                    "h2d_gbs_per_rank": h2d_gbs, "h2d_gbs_needed_at_value": r0["value"] / world_size * H * W / 1e9,
                    "host_cpus_near_gpu": len(near_cpus) if near_cpus else None},
            "gpu_launches": r0["launches"], "clocks": r0["clocks"], "roofline": roof, "kernels": kern, "cpu_baseline": cpu,
            "verified": all(all(d.values()) for d in all_ok), "verify": verified, "meta_verified": meta_ok,
        }
        for p in precisions[1:]:
            k2, c2 = kernel_table(res[p])
            line[f"value_{p}"] = res[p]["value"]
            line[f"ms_per_step_{p}"] = res[p]["ms"] / args.steps
            line[f"e2e_{p}"] = {"value": e2e[p]["value"], "unit": "frames/s", "h2d_bytes_per_step": e2e[p]["h2d"], "d2h_bytes_per_step": e2e[p]["d2h"]}
            line[f"kernels_{p}"] = {k: {"ms": v["ms"], "frac": v.get("frac")} for k, v in k2.items()}
            line[f"cnn_whole_{p}"] = c2
            line[f"dtype_{p}"] = DTYPE[p]
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=3, choices=[3, 4, 5], help="3: BASELINE configs[2] (the metric's; default), 4: 256 individuals / classes, "
                    "5: 3840x2160 + posture (outlines and midlines inside the step; headline roofline = the segmentation kernel vs HBM)")
    ap.add_argument("--batch", type=int, default=0, help="frames per step per GPU (default: 256 / 128 / 64 for config 3 / 4 / 5)")
    ap.add_argument("--slots", type=int, default=3, help="batches in flight in the end-to-end loop (handles + streams; the device-timed loop uses two)")
    ap.add_argument("--pool", type=int, default=4, help="distinct resident batches rotated through (defeats L2)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-frames", type=int, default=8, help="frames per step of the CPU arm / cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank to the CPUs near its GPU")
    ap.add_argument("--no-balance", action="store_true", help="end-to-end loop: equal frames per rank even when the ranks' H2D copy rates differ")
    ap.add_argument("--no-topo", action="store_true", help="local rank r uses CUDA device r (no PCIe-topology interleaving)")
    ap.add_argument("--no-pageable", action="store_true", help="skip the pageable-source e2e leg")
    ap.add_argument("--single-precision", action="store_true", help="measure only --precision")
    ap.add_argument("--precision", default="fp16c", choices=["fp16c", "bf16x3", "fp16"],
                    help="CNN arithmetic of the headline numbers: fp16c (fp16 MMA + one e5m2 correction MMA per k-step in conv2/conv3: 2 MMA slots; library "
                         "default; within 1e-3 on every weight set of tests/test_gpu_chain.py incl. logits of +-23), bf16x3 (3 MMAs per k-step, the most "
                         "accurate) or fp16 (1 MMA per k-step; within 1e-3 for O(1) logits only); the others are reported as value_<p>")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config], id=args.config)
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
