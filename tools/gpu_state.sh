#!/bin/bash
# State check on the GPU box: parity tests, full bench line, batch-1/2 streaming breakdown with and without the persistent FlowLM kernel. $1 = tag.
TAG=${1:-state}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_$TAG.json | cut -c1-1500
for B in 1 2; do
  B=$B python tools/batch1_quick.py 2>&1 | tail -1
  B=$B PTTS_B200_DEBUG_SKIP_MIMI=1 python tools/batch1_quick.py 2>&1 | tail -1
  B=$B PTTS_B200_PERSISTENT=0 python tools/batch1_quick.py 2>&1 | tail -1
done
