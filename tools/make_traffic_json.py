#!/usr/bin/env python3
"""profiles/attn_stream_traffic_<mode>.json from an `ncu --set full` capture that contains an attn_flow_split_kernel launch and the bench line
printed by the same (profiled) run; bench.py reads it for `roofline.traffic` of that mode (shared | private_kv | kv_f32).
Usage: python tools/make_traffic_json.py <capture.ncu-rep> <bench stdout log> <tag> [mode]"""
import csv
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def main(rep, log, tag, mode="shared"):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    val = [r for r in rows[2:] if "attn_flow_split_kernel" in r[hdr.index("Kernel Name")]][0]
    d = {h: (v, u) for h, u, v in zip(hdr, units, val)}
    line = [l for l in open(log) if l.startswith("{")][-1]
    b = json.loads(line)
    j = {"kernel": "attn_flow_split_kernel", "capture": tag,
         "dram_bytes_read": int(to_bytes(*d["dram__bytes_read.sum"])), "dram_bytes_write": int(to_bytes(*d["dram__bytes_write.sum"])),
         "gpu_time_us": float(d["gpu__time_duration.sum"][0].replace(",", "")) * {"us": 1, "ns": 1e-3, "ms": 1e3}[d["gpu__time_duration.sum"][1]],
         "algorithmic_bytes_at_capture": int(b["roofline"]["algorithmic_bytes_per_launch"]),
         "note": "one launch (one FlowLM layer, all utterances of the batch); ncu --set full --clock-control none, cold-cache replay"}
    json.dump(j, open(os.path.join(REPO, "profiles", f"attn_stream_traffic_{mode}.json"), "w"), indent=1)
    print(json.dumps(j))


if __name__ == "__main__":
    main(*sys.argv[1:5])
