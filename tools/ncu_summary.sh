#!/bin/bash
# key metrics of an .ncu-rep (read on the CPU box): tools/ncu_summary.sh gpurun_out/prof_q3.ncu-rep
for f in "$@"; do
  echo "== $f"
  ncu -i "$f" --page raw --csv 2>/dev/null | python3 -c '
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr,units,vals=rows[0],rows[1],rows[2]
want=["Kernel Name","Grid Size","Block Size","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","dram__throughput.avg.pct_of_peak_sustained_elapsed","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","lts__t_bytes.sum","l1tex__t_bytes.sum","sm__throughput.avg.pct_of_peak_sustained_elapsed","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","launch__occupancy_limit_warps","launch__shared_mem_per_block_dynamic","sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","smsp__inst_executed.sum","sm__cycles_elapsed.max","smsp__cycles_active.avg","lts__t_sector_hit_rate.pct","l1tex__t_sector_hit_rate.pct","smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct","smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct","smsp__warp_issue_stalled_barrier_per_warp_active.pct","smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct","smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct","smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct","smsp__warp_issue_stalled_wait_per_warp_active.pct","smsp__warp_issue_stalled_not_selected_per_warp_active.pct","smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct","smsp__warp_issue_stalled_membar_per_warp_active.pct","smsp__warp_issue_stalled_no_instruction_per_warp_active.pct","smsp__warp_issue_stalled_selected_per_warp_active.pct","smsp__warp_issue_stalled_drain_per_warp_active.pct","smsp__warp_issue_stalled_imc_miss_per_warp_active.pct","smsp__warp_issue_stalled_sleeping_per_warp_active.pct","smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct","smsp__warp_issue_stalled_tex_throttle_per_warp_active.pct"]
for w in want:
    for i,h in enumerate(hdr):
        if h==w: print(f"{h:90s} {vals[i]:>20s} {units[i]}")
'
done
