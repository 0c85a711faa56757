#!/usr/bin/env python3
"""profiles/attn_flow_traffic.json from an `ncu --set full` capture of one attn_flow_split_kernel launch and the bench line printed by the
same (profiled) run. Usage: python tools/make_traffic_json.py <capture.ncu-rep> <bench stdout log> <tag>"""
import csv
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def main(rep, log, tag):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, val = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, val)}
    assert "attn_flow_split_kernel" in d["Kernel Name"][0], d["Kernel Name"][0]
    line = [l for l in open(log) if l.startswith("{")][-1]
    b = json.loads(line)
    j = {"kernel": "attn_flow_split_kernel", "capture": tag,
         "dram_bytes_read": int(to_bytes(*d["dram__bytes_read.sum"])), "dram_bytes_write": int(to_bytes(*d["dram__bytes_write.sum"])),
         "gpu_time_us": float(d["gpu__time_duration.sum"][0].replace(",", "")) * {"us": 1, "ns": 1e-3, "ms": 1e3}[d["gpu__time_duration.sum"][1]],
         "algorithmic_bytes_at_capture": int(b["roofline"]["algorithmic_bytes_per_launch"]),
         "note": "one launch (one FlowLM layer, all utterances of the batch); ncu --set full --clock-control none, cold-cache replay"}
    json.dump(j, open(os.path.join(REPO, "profiles", "attn_flow_traffic.json"), "w"), indent=1)
    print(json.dumps(j))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3])
