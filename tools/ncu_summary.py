#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the handful of numbers the roofline discussion needs.
usage: python tools/ncu_summary.py report.ncu-rep [> profiles/rNN_<kernel>.txt]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "derived__lts__lts2xbar_bytes.sum.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_op_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "launch__cluster_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        print(f"kernel: {d.get('Kernel Name', '')[:110]}")
        print(f"grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d and d[k] not in ("", "n/a"):
                print(f"  {k:78s} {d[k]:>16s} {u.get(k, '')}")
        for k in hdr:  # every tensor-pipe activity metric the report has (imma for the INT8 GEMMs, hmma for the fp16 attention)
            if "pipe_tensor" in k and k not in KEYS and d[k] not in ("", "n/a", "0"):
                print(f"  {k:78s} {d[k]:>16s} {u.get(k, '')}")
        stalls = sorted(((float(d[k]), k[len(STALL):].replace("_per_issue_active.ratio", "")) for k in hdr
                         if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and d[k] not in ("", "n/a")), reverse=True)
        print("  warp stall reasons (warps stalled per issue-active cycle):")
        for v, k in stalls[:8]:
            print(f"    {k:40s} {v:8.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
