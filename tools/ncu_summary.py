#!/usr/bin/env python
"""Development aid: one line per profiled launch of an .ncu-rep (ncu --set full), with the metrics the design notes quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>_summary.txt
"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("launch__shared_mem_per_block_dynamic", "dyn smem KB"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"), ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "hmma %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %")]

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
print(f"# {rep}: ncu --set full --clock-control none (cold-cache, serialised launches: compare shares, not absolutes)")
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    name = name.replace("void <unnamed>::", "")[:60]
    parts = []
    for key, label in WANT:
        if key in idx and r[idx[key]] != "":
            u = units[idx[key]]
            parts.append(f"{label} {r[idx[key]]}{(' ' + u) if u and u not in ('%',) and label in ('dram read', 'dram write', 'us') else ''}")
    print(f"{name}\n    " + "; ".join(parts))
