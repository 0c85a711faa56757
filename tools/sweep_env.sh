#!/bin/bash
# Quick A/B of engine tuning hooks on the bench workload (primary mode only): each argument is "NAME=VALUE[,NAME=VALUE...]" or "-" for the defaults.
for cfg in "$@"; do
  envs=$(echo "$cfg" | tr ',' ' ')
  [ "$cfg" = "-" ] && envs=""
  out=$(env $envs timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline $SWEEP_ARGS 2>/dev/null | tail -1)
  python - "$cfg" <<PY "$out"
import json, sys
d = json.loads(sys.argv[2])
r = d.get("roofline", {})
print("SWEEP", sys.argv[1], "value", d["value"], "ms", d["ms_per_step"], "e2e", d.get("e2e", {}).get("value"), "stream_ms", r.get("avg_launch_ms"), "tile_ms", r.get("prefix_tiles", {}).get("avg_launch_ms"), "verify", {k: d["verify"][k] for k in ("latent_maxabs", "latent_rel", "snr_db_min")} if "verify" in d else None)
PY
done
