#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof_tc.ncu-rep > profiles/r01_ncu_gemm_tc.txt"""
import csv
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor", "sm__pipe_tensor", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit",
        "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "smsp__warp_issue_stalled", "gpc__cycles_elapsed.max", "smsp__average_warp")


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        print("# kernel:", r[head.index("Kernel Name")])
        print("# source: %s (ncu --set full --clock-control none, one launch; cold-cache, serialised)" % rep)
        for i, name in enumerate(head):
            if any(name.startswith(k) for k in KEEP) and r[i] not in ("", "n/a"):
                print("%-78s %18s %s" % (name, r[i], units[i]))
        rd = wr = None
        for i, name in enumerate(head):
            if name == "dram__bytes_read.sum": rd = (float(r[i]), units[i])
            if name == "dram__bytes_write.sum": wr = (float(r[i]), units[i])
        if rd and wr:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            print("%-78s %18.1f MB" % ("DRAM traffic (read + write) per launch", (rd[0] * scale[rd[1]] + wr[0] * scale[wr[1]]) / 1e6))


if __name__ == "__main__":
    main()
