#!/usr/bin/env python
"""ncu-rep -> compact JSON summary (all raw metrics of the first captured kernel) for profiles/.
    python scripts/tools/ncu_summary.py gpurun_out/mas_full.ncu-rep profiles/r1_ncu_full_xxx.json"""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: [v, u] for h, u, v in zip(hdr, units, vals)}
json.dump(d, open(out, "w"), indent=1)
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
for k in keys:
    for h in d:
        if h == k or h.endswith(k):
            print(f"{h}: {d[h][0]} {d[h][1]}")
            break
