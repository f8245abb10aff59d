#!/usr/bin/env python
"""Markdown table of the key metrics of several .ncu-rep files (read on the CPU box):
   python tools/ncu_table.py label1=path1.ncu-rep label2=path2.ncu-rep ..."""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]
cols = []
for arg in sys.argv[1:]:
    label, path = arg.split("=", 1)
    if path.endswith(".csv"):  # already exported with `ncu -i x.ncu-rep --page raw --csv`
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    cols.append((label, d))
print("| metric | " + " | ".join(c[0] for c in cols) + " |")
print("|---|" + "---|" * len(cols))
print("| kernel | " + " | ".join("`" + c[1].get("Kernel Name", ("?", ""))[0][:70] + "`" for c in cols) + " |")
for w in WANT:
    cells = []
    for _, d in cols:
        v, u = d.get(w, ("n/a", ""))
        try:
            v = f"{float(v.replace(',', '')):.6g}"
        except ValueError:
            pass
        cells.append(f"{v} {u}".strip())
    print(f"| `{w}` | " + " | ".join(cells) + " |")
