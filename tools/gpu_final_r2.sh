#!/bin/bash
# Final bench lines of the round: full bench.py line (roofline, e2e, verify, sustained, extras, cpu_baseline), the other BASELINE configs, batch-1/2 breakdown.
TAG=${1:-r2}
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
timeout 900 python tools/bench_configs.py --skip-cpu > gpurun_out/bench_configs_$TAG.jsonl 2> gpurun_out/bench_configs_$TAG.err; echo "configs rc=$?"
for B in 1 2; do B=$B python tools/batch1_quick.py 2>&1 | tail -1; done > gpurun_out/batch1_$TAG.txt
cat gpurun_out/batch1_$TAG.txt; tail -5 gpurun_out/bench_configs_$TAG.jsonl | cut -c1-300
