#!/usr/bin/env python3
"""Key metrics of every kernel in an .ncu-rep (read on the CPU box): duration, DRAM bytes, DRAM/L2/SM/tensor utilisation, occupancy, registers."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__cycles_active.avg",
        "lts__t_sector_hit_rate.pct", "launch__waves_per_multiprocessor"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")][:60]
        print(f"== {name}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"   {w:72s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(f"# {p}")
        main(p)
