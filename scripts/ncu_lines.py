#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep (ncu --page source --csv)."""
import csv, linecache, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = None; hdr = None; agg = {}; smp = {}
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1]; continue
    if len(r) == 2: continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < 8: continue
    try: ln = int(r[0])
    except ValueError: continue
    k = (cur, ln)
    def num(x):
        try: return float(x)
        except ValueError: return 0.0
    agg[k] = agg.get(k, 0) + num(r[hdr.index('Instructions Executed')])
    smp[k] = smp.get(k, 0) + num(r[hdr.index('# Samples')])
ti = sum(agg.values()) or 1; ts = sum(smp.values()) or 1
print(f"total warp instructions {ti:.0f}, samples {ts:.0f}")
for k, v in sorted(agg.items(), key=lambda kv: -(kv[1] / ti + smp[kv[0]] / ts))[:top]:
    print(f"{k[0].split('/')[-1]}:{k[1]:5d} {v / ti * 100:5.1f}% inst {smp[k] / ts * 100:5.1f}% samp | {linecache.getline(k[0], k[1]).strip()[:100]}")
