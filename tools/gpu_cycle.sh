#!/bin/bash
# One GPU iteration: parity tests, bench line, ncu launch list of two timed steps (tag = $1), optional full capture ($2 = "-k ... -s .. -c .." args).
TAG=${1:-x}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print("BENCH", d["value"], d["ms_per_step"], "e2e", d.get("e2e", {}).get("value"), d["roofline"]["frac"], d["roofline"]["segments_ms_per_step"])
PY
PTTS_NCU_RANGE=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$TAG.log 2>&1
if [ -n "$2" ]; then
  PTTS_NCU_RANGE=1 timeout 400 ncu --profile-from-start off --set full --import-source on --clock-control none $2 -f -o gpurun_out/full_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_full_$TAG.log | cut -c1-200
fi
