#!/usr/bin/env python
"""Compact text summary of an `ncu --set full` report (one kernel): duration, DRAM traffic, pipe / memory utilisation,
stall mix.  ncu_summary.py file.ncu-rep > profiles/xxx.summary.txt"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, v = rows[0], rows[1], rows[-1]
d = {n: (val, unit) for n, unit, val in zip(h, u, v)}
def g(n):
    return d.get(n, ("n/a", ""))
print("kernel      :", g("Kernel Name")[0])
print("grid/block  :", g("Grid Size")[0], "/", g("Block Size")[0], " regs/thread", g("launch__registers_per_thread")[0],
      " dyn smem/block", g("launch__shared_mem_per_block_dynamic")[0], g("launch__shared_mem_per_block_dynamic")[1])
keys = [
 ("gpu__time_duration.sum", "duration"),
 ("sm__cycles_elapsed.avg.per_second", "SM clock during capture"),
 ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
 ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
 ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
 ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
 ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput % of peak"),
 ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "tensor-core smem operand wavefronts % of peak"),
 ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
 ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "tensor subpipe (hmma) % of peak"),
 ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
 ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
 ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
 ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
 ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
 ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
 ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
 ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected / issue"),
 ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction / issue"),
]
for k, label in keys:
    val, unit = g(k)
    print(f"{label:52s}: {val} {unit}")
