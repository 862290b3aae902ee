"""Pick the metrics the round summaries quote out of an `ncu --page raw --csv` export."""
import csv, sys
WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size"]
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        print(path, "empty"); continue
    h, u, v = rows[0], rows[1], rows[2]
    print("==", path, v[h.index("Kernel Name")][:60])
    for w in WANT:
        if w in h:
            i = h.index(w)
            print(f"  {w:70s} {v[i]:>16s} {u[i]}")
