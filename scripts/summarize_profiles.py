#!/usr/bin/env python
"""Markdown table of the round's bench lines: profiles/r2_bench_cfg{3,4,5}_{1,2,4,8}gpu.json -> stdout (profiles/r2_summary.md)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
print("| config | GPUs | precision | frames/s device-timed | ms/step | frames/s e2e | H2D GB/s per rank | dominant kernel: frac of measured peak | verified | also measured (device / e2e) |")
print("|---|---|---|---|---|---|---|---|---|---|")
for c in (3, 4, 5):
    for n in (1, 2, 4, 8):
        p = os.path.join(ROOT, "profiles", f"r2_bench_cfg{c}_{n}gpu.json")
        if not os.path.exists(p):
            continue
        d = json.loads(open(p).read().strip().splitlines()[-1])
        r = d["roofline"]
        others = "; ".join(f"{q}: {d['value_' + q]:,.0f} / {d['e2e_' + q]['value']:,.0f}" for q in ("fp16c", "bf16x3", "fp16") if f"value_{q}" in d)
        h2d = ", ".join(f"{g[1]:.0f}" for g in d["e2e"]["h2d_gbs_per_rank"])
        print(f"| {c} | {n} | {d['precision']} | {d['value']:,.0f} | {d['ms_per_step']:.2f} | {d['e2e']['value']:,.0f} | {h2d} | {r['kernel']}: {r['frac']:.2f} ({r['bound']}) | "
              f"{d['verified']}{'' if d.get('meta_verified') is None else ' / meta ' + str(d['meta_verified'])} | {others} |")
