#!/bin/bash
# bench only (no tests, no ncu): prints value / ms_per_step / segments. Extra env via "VAR=.. VAR=.." in $1.
env $1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('BENCH $1', d['value'], d['ms_per_step'], d['roofline']['segments_ms_per_step'])"
