#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the handful of numbers the roofline
discussion uses: duration, DRAM bytes, throughput percentages, issue utilisation, stall mix."""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        kname = r[head.index("Kernel Name")]
        print(f"kernel: {kname}")
        for w in WANT:
            if w in head:
                i = head.index(w)
                print(f"  {w:88s} {r[i]:>16s} {units[i]}")
        stalls = sorted(((float(r[i]), h) for i, h in enumerate(head)
                         if "average_warps_issue_stalled" in h and "not_issued" not in h and r[i]), reverse=True)
        print("  warp stall reasons (warps per issue-active cycle):")
        for v, h in stalls[:8]:
            print(f"    {h.split('issue_stalled_')[1].split('_per_')[0]:24s} {v:.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
