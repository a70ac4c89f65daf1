"""`ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py > profiles/NAME.csv`
Keeps the columns of a `--set full` capture that the profiles/ README reads (time, DRAM bytes, hit rates, occupancy,
instructions, issue rate, FP64 pipe, stall ratios): one row per captured launch, units in the second row."""
import csv
import sys

COLS = ["ID", "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]

rows = [r for r in csv.reader(sys.stdin) if r]
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
head, units, body = rows[hi], rows[hi + 1], rows[hi + 2:]
idx = [head.index(c) for c in COLS if c in head]
w = csv.writer(sys.stdout)
w.writerow([head[i] for i in idx])
w.writerow([units[i] for i in idx])
for r in body:
    if len(r) > max(idx):
        w.writerow([r[i] for i in idx])
