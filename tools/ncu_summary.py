"""Summarise an .ncu-rep (raw page) into the handful of counters the design cares about.
usage: python tools/ncu_summary.py file.ncu-rep [more.ncu-rep ...]"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__warps_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warps_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warps_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warps_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warps_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warps_issue_stalled_wait_per_warp_active.pct",
        "smsp__warps_issue_stalled_not_selected_per_warp_active.pct"]

for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        rec = dict(zip(hdr, vals))
        print(f"== {path}: {rec.get('Kernel Name', '?')[:90]}")
        for h, u in zip(hdr, units):
            if h in WANT:
                print(f"   {h:78s} {rec[h]:>16s} {u}")
