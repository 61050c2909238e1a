#!/usr/bin/env python
"""Condense an Nsight Compute report into the few numbers DESIGN.md / profiles/ quote.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md
"""
import csv
import subprocess
import sys

KEYS = [
    ("time [ms]", "gpu__time_duration.sum"),
    ("regs/thread", "launch__registers_per_thread"),
    ("DRAM read", "dram__bytes_read.sum"),
    ("DRAM write", "dram__bytes_write.sum"),
    ("DRAM % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warps active % of peak", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("fp64 pipe busy %", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    ("warp instructions", "smsp__inst_executed.sum"),
    ("active lanes / instr", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("L1 hit %", "l1tex__t_sector_hit_rate.pct"),
    ("L2 hit %", "lts__t_sector_hit_rate.pct"),
    ("stall long_scoreboard / issue", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("stall wait / issue", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("stall math_pipe / issue", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
    ("stall lg_throttle / issue", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    print("# %s\n" % path.split("/")[-1])
    print("ncu --set full --clock-control none (one launch each, ~40 replays: cold-cache, serialised; "
          "compare shares, not absolutes)\n")
    for r in rows[2:]:
        name = r[head.index("Kernel Name")]
        grid, block = r[head.index("Grid Size")], r[head.index("Block Size")]
        print("## `%s`\n\ngrid %s, block %s\n\n| metric | value |\n|---|---|" % (name, grid, block))
        for label, key in KEYS:
            if key in head:
                i = head.index(key)
                print("| %s | %s %s |" % (label, r[i], units[i]))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
