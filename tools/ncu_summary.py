#!/usr/bin/env python3
"""Condenses an .ncu-rep (ncu --set full) into the handful of counters the roofline argument uses.
usage: tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [more.ncu-rep ...] > profiles/xxx.txt"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__issue_active.avg.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_sleeping_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct",
        "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct"]
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"== {rep}: {d.get('Kernel Name','?')[:60]}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d:
                print(f"   {k:75s} {d[k]:>18s} {units[hdr.index(k)]}")
        hits = [h for h in hdr if "fp64" in h and h not in KEYS]
        for h in hits[:8]:
            print(f"   {h:75s} {d[h]:>18s} {units[hdr.index(h)]}")
