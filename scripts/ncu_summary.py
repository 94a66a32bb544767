"""Per-kernel summary of an ncu report (raw page): usage  ncu -i rep --page raw --csv | python scripts/ncu_summary.py"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
ki = hdr.index("Kernel Name")
units = rows[1]
for r in rows[2:]:
    print("==", r[ki].split("(")[0])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"   {w:100s} {r[i]:>14s} {units[i]}")
