#!/bin/bash
# Sentence-start time and parity of the prefill tile attention at operand precision $1 (2 = q and P as hi+lo bf16 pairs, 1 = q only, 0 = plain bf16)
export PTTS_B200_PREFILL_PREC=$1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --verify 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PREFILL_PREC $1 sentence_start_ms', d['config']['sentence_start_ms'], 'verify', {k: d['verify'][k] for k in ('pass','latent_maxabs','latent_rel','snr_db_min','kv_maxabs')})"
timeout 600 python -m pytest tests/test_gpu_bench_config.py tests/test_gpu_ref_golden.py -x -q 2>&1 | tail -1
