#!/usr/bin/env python3
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: count, total time, share."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0][:70]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(row["Metric Unit"], 1.0)
        agg[name][0] += 1; agg[name][1] += v; tot += v
    print(f"# {path}: per-kernel device time (cold-cache, serialised under ncu: compare SHARES, not absolutes)")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:72s} n={n:5d} total={t / 1e3:11.1f} us  share={t / tot * 100:5.1f}%")
    print(f"total {tot / 1e3:.1f} us over {sum(n for n, _ in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1])
