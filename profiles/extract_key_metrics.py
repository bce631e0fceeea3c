"""Reduce `ncu -i X.ncu-rep --page raw --csv` to the columns the roofline discussion uses (one row per launch)."""
import csv
import sys

KEEP = ["ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__inst_executed.sum"]
rows = list(csv.reader(open(sys.argv[1], newline="")))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
idx = [h.index(k) for k in KEEP if k in h]
w = csv.writer(open(sys.argv[2], "w", newline=""))
for r in rows[hi:]:
    if len(r) == len(h):
        row = [r[i] for i in idx]
        row[1] = row[1].replace("gccvae::", "")[:70]
        w.writerow(row)
