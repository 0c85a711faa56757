#!/bin/bash
# Quick A/B of engine tuning hooks on the bench workload (primary mode only): each argument is "NAME=VALUE[,NAME=VALUE...]" or "-" for the defaults.
for cfg in "$@"; do
  envs=$(echo "$cfg" | tr ',' ' ')
  [ "$cfg" = "-" ] && envs=""
  out=$(env $envs timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | tail -1)
  python - "$cfg" <<PY "$out"
import json, sys
d = json.loads(sys.argv[2])
r = d.get("roofline", {})
print("SWEEP", sys.argv[1], "value", d["value"], "ms", d["ms_per_step"], "e2e", d.get("e2e", {}).get("value"), "stream_ms", r.get("avg_launch_ms"), "seg", r.get("segments_ms_per_step"))
PY
done
