#!/bin/bash
# Round evidence on ONE B200 (run through gpurun from the repo root): the GPU parity suite, memcheck over the posture / recount tests, the three
# BASELINE configs at N = 1, the reference arm, and the micro-benchmarks.
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gpu_head.txt
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_midline.py tests/test_gpu_outline.py -x -q -m gpu -k "not benchmark_workload" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid" | tail -5 | tee gpurun_out/r2_memcheck_posture.txt
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_seg.py -x -q -m gpu -k "recount" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid" | tail -3 | tee -a gpurun_out/r2_memcheck_posture.txt
timeout 500 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_vi.py -x -q -m gpu -k "reference_class_golden or 256_classes" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid" | tail -3 | tee gpurun_out/r2_memcheck_cnn.txt
timeout 500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_vi.py -x -q -m gpu -k "reference_class_golden and fp16" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard" | tail -4 | tee -a gpurun_out/r2_memcheck_cnn.txt
for c in 3 4 5; do timeout 400 python bench.py --config $c > gpurun_out/r2_cfg${c}_1gpu.json 2> gpurun_out/r2_cfg${c}_1gpu.err; tail -c 300 gpurun_out/r2_cfg${c}_1gpu.err; done
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null
timeout 100 python scripts/bench_posture.py 128 1920x1080 10 > gpurun_out/r2_posture_micro.jsonl; timeout 100 python scripts/bench_posture.py 64 3840x2160 10 >> gpurun_out/r2_posture_micro.jsonl
python - <<PY
import json
for c in (3, 4, 5):
    d = json.loads(open("gpurun_out/r2_cfg%d_1gpu.json" % c).read().strip().splitlines()[-1])
    print("config", c, round(d["value"]), round(d["e2e"]["value"]), "verified", d["verified"], round(d["roofline"]["frac"], 3), round(d["kernels"]["seg_rle"]["frac"], 3), d["cpu_baseline"]["value"] if d["cpu_baseline"] else None, d["clocks"]["reasons"])
r = json.loads(open("gpurun_out/r2_bench_reference_arm.json").read().strip().splitlines()[-1]); print("reference arm", r["value"], r.get("cpu_baseline", {}).get("cores"))
PY
cat gpurun_out/r2_posture_micro.jsonl | cut -c1-250
