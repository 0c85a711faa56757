#!/bin/bash
# Batch sweep of the two pipeline forms: 7 + 6 graphs interleaved (default above PTTS_B200_IMMEDIATE_BELOW) vs one graph per stream.
for B in 32 64 128 192 256; do
  for IB in 32 100000; do
    PTTS_B200_IMMEDIATE_BELOW=$IB timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --batch $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SWEEP batch $B immediate_below $IB', d['value'], d['ms_per_step'])"
  done
done
