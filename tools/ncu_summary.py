#!/usr/bin/env python
"""Turn an .ncu-rep (one kernel launch, `ncu --set full --import-source on`) into the markdown
summary committed under profiles/.   Usage: python tools/ncu_summary.py X.ncu-rep [title] > profiles/Y.md"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
m = {k: (v[i], u[i]) for i, k in enumerate(h)}
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "dram__bytes_write.sum.per_second", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_srcunit_tex_op_atom.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_write.sum"]
print(f"# {title}\n")
print(f"source: `{rep}` (ncu --set full --clock-control none --import-source on; one launch, cold cache, serialised)\n")
print("| metric | value | unit |\n|---|---|---|")
for k in want:
    if k in m:
        print(f"| {k} | {m[k][0]} | {m[k][1]} |")
try:
    rd = float(m["dram__bytes_read.sum"][0]); wr = float(m["dram__bytes_write.sum"][0])
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}
    tot = rd * scale[m["dram__bytes_read.sum"][1]] + wr * scale[m["dram__bytes_write.sum"][1]]
    print(f"\nDRAM traffic of this launch (read + write): **{tot/1e9:.3f} GB**\n")
except Exception:
    pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
top = subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_top.py"), "24"], input=src, capture_output=True, text=True).stdout
print("## Warp-stall sampling (SASS, top instructions)\n\n```\n" + top + "```")
