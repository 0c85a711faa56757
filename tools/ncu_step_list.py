#!/usr/bin/env python3
"""Per-launch list of ONE step from an ncu launch list (gpu__time_duration.sum --csv): index, kernel, grid, block, microseconds."""
import csv
import sys


def main(path, steps=2):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r["Metric Name"] == "gpu__time_duration.sum"]
    n = len(rows) // steps
    tot = 0.0
    for i, r in enumerate(rows[:n]):
        v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[r["Metric Unit"]]
        tot += v
        print(i, r["Kernel Name"].split("(")[0][:44], r["Grid Size"], r["Block Size"], f"{v:.1f}")
    print(f"one step: {n} launches, {tot:.1f} us (cold-cache, serialised)")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2)
