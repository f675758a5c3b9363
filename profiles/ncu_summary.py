#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): per-kernel duration, DRAM
traffic, throughput percentages and the top warp-stall reasons.
usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep [> profiles/x.txt]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_global_atom.sum",
        "sm__cycles_elapsed.max"]
for r in rows[2:]:
    print("=" * 100)
    print(r[idx["Kernel Name"]][:140])
    for k in keys:
        if k in idx:
            print(f"  {k:70s} {r[idx[k]]:>18s} {units[idx[k]]}")
    st = []
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                st.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace(
                    "_per_issue_active.ratio", "")))
            except ValueError:
                pass
    print("  warp stalls per issue (top): " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:7]))
