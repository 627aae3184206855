#!/usr/bin/env python3
"""Kernel summary for profiles/: the metrics bench.py's roofline.pipe_utilisation reads, from
`ncu -i X.ncu-rep --page raw --csv`.

usage: ncu -i X.ncu-rep --page raw --csv > raw.csv
       python tools/ncu_summary.py raw.csv "<headline describing kernel / workload / command>" > profiles/rNN_ncu_render_kernel_summary.txt"""
import csv
import sys

KEEP = """dram__bytes_read.sum dram__bytes_write.sum gpu__time_duration.sum launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem launch__registers_per_thread launch__grid_size launch__block_size
sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed
sm__throughput.avg.pct_of_peak_sustained_elapsed sm__warps_active.avg.per_cycle_active
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active smsp__sass_average_branch_targets_threads_uniform.pct
smsp__thread_inst_executed_per_inst_executed.ratio smsp__warps_eligible.avg.per_cycle_active
smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed
smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed
smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed
smsp__sass_thread_inst_executed_op_dfma_pred_on.sum smsp__sass_thread_inst_executed_op_dmul_pred_on.sum
smsp__sass_thread_inst_executed_op_dadd_pred_on.sum sm__cycles_elapsed.avg sm__cycles_elapsed.max""".split()


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print(sys.argv[2] if len(sys.argv) > 2 else "ncu --set full --clock-control none")
    col = {h: k for k, h in enumerate(hdr)}
    for name in KEEP:
        if name in col:
            k = col[name]
            print("%s [%s] = %s" % (name, units[k], vals[k]))


if __name__ == "__main__":
    main()
