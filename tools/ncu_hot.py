#!/usr/bin/env python3
"""Top SASS instructions by warp-stall samples from an .ncu-rep source page, with the dominant stall reason."""
import csv
import subprocess
import sys


def main(path, top=25, skip=0):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--launch-skip", str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    isrc, isamp = hdr.index("Source"), hdr.index("# Samples")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    total = 0
    for r in rows[2:]:
        if len(r) <= isamp:
            continue
        try:
            n = int(r[isamp])
        except ValueError:
            continue
        total += n
        if n:
            st = sorted(((int(r[i]) if r[i].isdigit() else 0, h) for i, h in stall_cols), reverse=True)[:2]
            data.append((n, r[isrc][:110], st))
    data.sort(reverse=True)
    print(f"# {path} (launch {skip}: {rows[0][1][:60]}): total samples {total}")
    for n, src, st in data[:top]:
        print(f"{n:7d} {100.0 * n / max(total, 1):5.1f}%  {src:110s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
