#!/usr/bin/env python
"""Text summary of an ncu report: selected raw metrics per kernel + the hottest source lines.
    python profiles/scripts/ncu_summary.py REPORT.ncu-rep KERNEL_REGEX [N_LINES]"""
import csv
import io
import os
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = sys.argv[3] if len(sys.argv) > 3 else "25"
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + kern],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("== %s" % r[ix["Kernel Name"]])
        for k in KEEP:
            if k in ix:
                print("   %-90s %s %s" % (k, r[ix[k]], units[ix[k]]))
    here = os.path.dirname(os.path.abspath(__file__))
    print("-- hottest source lines (samples / warp instructions)")
    sys.stdout.flush()
    subprocess.run([sys.executable, os.path.join(here, "..", "ncu_top_lines.py"), rep, kern, top])


if __name__ == "__main__":
    main()
