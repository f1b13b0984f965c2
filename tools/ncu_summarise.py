#!/usr/bin/env python
"""Compact per-launch summary of an .ncu-rep: `ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summarise.py > out.csv`.

Keeps the metrics the profiles/ summaries quote (duration, DRAM bytes, issue activity, pipe utilisation, occupancy,
registers, stall reasons), one column per captured launch.
"""
import csv
import sys

KEEP = [
    "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def main():
    rows = list(csv.reader(line for line in sys.stdin if line.startswith('"')))
    header, units, launches = rows[0], rows[1], rows[2:]
    col = {name: i for i, name in enumerate(header)}
    out = csv.writer(sys.stdout)
    out.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(launches))])
    for name in KEEP:
        if name in col:
            out.writerow([name, units[col[name]]] + [r[col[name]] for r in launches])


if __name__ == "__main__":
    main()
