#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` report (first instance of every kernel name): duration, DRAM bytes,
DRAM / SM throughput %, tensor-pipe activity, IPC, occupancy, registers, grid.  Usage: ncu_summary.py x.ncu-rep"""
import csv, io, subprocess, sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
        ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
        ("sm__inst_executed.avg.per_cycle_active", "ipc"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("lts__t_bytes.sum", "l2_bytes")]
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
seen = {}
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("tb::", "").replace("tc::", "")
    if name in seen:
        seen[name][1] += 1
        continue
    vals = []
    for m, short in WANT:
        if m in hdr:
            i = hdr.index(m)
            vals.append(f"{short}={r[i]}{(' ' + units[i]) if units[i] and units[i] not in ('%',) else ''}")
    seen[name] = [vals, 1]
for name, (vals, n) in seen.items():
    print(f"{name}  (x{n} in the capture)")
    print("    " + ", ".join(vals))
