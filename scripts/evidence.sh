#!/bin/bash
# Round evidence on a B200 box (run through gpurun from the repo root): GPU parity suite, both bench arms, network micro-benchmark.
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_head.log
timeout 280 python bench.py > gpurun_out/bench_head.json 2>gpurun_out/bench_head.err; tail -c 300 gpurun_out/bench_head.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_head.json 2>/dev/null
timeout 200 python scripts/bench_nets.py 2048 5 2>&1 | grep "^{" > gpurun_out/nets.jsonl
timeout 150 python scripts/bench_seg.py 128 10 1 gray none 1920x1080 outlines 2>&1 | grep "^{" > gpurun_out/seg_outlines.jsonl
timeout 150 python scripts/bench_seg.py 32 10 1 gray none 3840x2160 outlines 2>&1 | grep "^{" >> gpurun_out/seg_outlines.jsonl
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_posture.csv python scripts/bench_seg.py 128 2 1 gray none 1920x1080 outlines > /dev/null 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_head.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), d["roofline"]["frac"], d["kernels"]["seg_rle"]["frac"], d["cpu_baseline"]["value"], d["clocks"])
r = json.loads(open("gpurun_out/bench_ref_head.json").read().strip().splitlines()[-1]); print(r["value"], r.get("cpu_baseline"))
for l in open("gpurun_out/seg_outlines.jsonl"):
    n = json.loads(l); print(n["size"], n["B"], n["blobs"], "outlines ms", n["outlines_ms_incl_d2h"], "midlines ms", n["midlines_ms_incl_d2h"])
for l in open("gpurun_out/nets.jsonl"):
    n = json.loads(l); print(n["version"], n["precision"], round(n["crops_per_s"]), round(n["tflops_algorithmic"], 1))
PY
