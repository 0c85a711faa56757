#!/bin/bash
# bench only (no tests, no ncu): prints value / ms_per_step / segments. $1 = extra bench.py args, $2 = extra env ("VAR=.. VAR=..").
env $2 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e $1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('BENCH [$1 $2]', d['value'], d['ms_per_step'], d['roofline']['segments_ms_per_step'])"
