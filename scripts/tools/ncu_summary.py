#!/usr/bin/env python
"""ncu-rep -> compact JSON summary (the raw metrics of the first captured kernel that the docs / bench.py cite; pass
--all as third argument for every metric) for profiles/.
    python scripts/tools/ncu_summary.py gpurun_out/mas_full.ncu-rep profiles/r1_ncu_full_xxx.json"""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: [v, u] for h, u, v in zip(hdr, units, vals)}
import re
KEEP = re.compile(r"^(Kernel Name|Block Size|Grid Size|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|dram__cycles_active\.avg|gpu__dram_throughput\..*pct.*elapsed|sm__throughput\..*pct.*elapsed|"
                  r"launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic|occupancy_limit_.*|waves_per_multiprocessor)|"
                  r"sm__warps_active\.avg\.pct_of_peak_sustained_active|sm__inst_executed\.sum|sm__inst_issued\.sum|smsp__issue_active\.avg\.pct.*|smsp__inst_executed\.sum|"
                  r"sm__pipe_tensor.*cycles_active.*pct.*|sm__inst_executed_pipe_tensor.*sum|sm__ops_path_tensor.*|smsp__average_warps?_issue_stalled_.*_per_issue_active\.ratio|"
                  r"smsp__average_warp_latency_issue_stalled_.*|l1tex__data_bank_conflicts_pipe_lsu.*sum|l1tex__data_pipe_lsu_wavefronts_mem_shared.*sum|"
                  r"lts__t_bytes\.sum|lts__t_sectors_srcunit_tex_op_(read|write)\.sum|sm__cycles_elapsed\.(max|avg)|smsp__cycles_active\.avg|sm__cycles_active\.avg)$")
def keep(k, v):
    if not KEEP.match(k):
        return False
    if k.startswith(("sm__ops_path_tensor", "sm__inst_executed_pipe_tensor", "sm__pipe_tensor")):      # only what is non-zero
        try:
            nz = float(str(v[0]).replace(",", "")) != 0
        except ValueError:
            nz = False
        return nz and ".min" not in k and ".max" not in k and k.endswith((".sum", "pct_of_peak_sustained_active", "pct_of_peak_sustained_elapsed"))
    return True
json.dump(d if len(sys.argv) > 3 and sys.argv[3] == "--all" else {k: v for k, v in d.items() if keep(k, v)}, open(out, "w"), indent=0)
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
for k in keys:
    for h in d:
        if h == k or h.endswith(k):
            print(f"{h}: {d[h][0]} {d[h][1]}")
            break
