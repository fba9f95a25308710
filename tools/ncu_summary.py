#!/usr/bin/env python
"""Summarise an ncu report: `ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py`.
Prints, per profiled launch, the handful of metrics the roofline discussion uses."""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_uniform.sum",
        "smsp__inst_executed.sum", "launch__shared_mem_per_block_dynamic"]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
for row in rows[2:]:
    d = dict(zip(hdr, row))
    print("kernel:", d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for i, h in enumerate(hdr):
        if h in WANT:
            print(f"  {h} = {row[i]} {units[i]}")
