#!/usr/bin/env python
"""Summarise an ncu report (read here, no GPU needed) into the text kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<round>_<what>.txt
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, check=True).stdout.decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print("# %s  (ncu --set full --clock-control none; values per launch)" % rep)
    for r in rows[2:]:
        if r[col["sm__cycles_elapsed.avg"]] in ("", "-nan", "nan"):
            continue      # a launch whose replay passes were cut short
        print("\n== launch %s  %s  grid %s block %s" % (r[col["ID"]], r[col["Kernel Name"]],
                                                      r[col["Grid Size"]], r[col["Block Size"]]))
        for m in METRICS:
            if m in col and r[col[m]] not in ("", "n/a"):
                print("  %-86s %s %s" % (m, r[col[m]], units[col[m]]))


if __name__ == "__main__":
    main()
