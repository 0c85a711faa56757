#!/bin/bash
# Round-2 evidence run on the GPU box: plain bench (must exit 0) -> ncu launch list of two timed steps -> ncu --set full of the FlowLM
# attention kernels (shared-prefix tile kernel, per-utterance stream kernel, merge) and of the largest SEANet GEMM. $1 = tag.
TAG=${1:-r2}
ARGS="--steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
python bench.py $ARGS > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
PTTS_NCU_RANGE=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py $ARGS > gpurun_out/ncu_$TAG.log 2>&1
PTTS_NCU_RANGE=1 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"attn_(tile|flow_split|merge)" -c 6 -f -o gpurun_out/full_attn_$TAG \
  python bench.py $ARGS > gpurun_out/ncu_full_attn_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_attn_$TAG.log | cut -c1-300
PTTS_NCU_RANGE=1 timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:"gemm_tc_kernel<128, 2>|gemm_tc_kernel<64, 3>|attn_mimi" -c 4 -f -o gpurun_out/full_mimi_$TAG \
  python bench.py $ARGS > gpurun_out/ncu_full_mimi_$TAG.log 2>&1
ls -la gpurun_out/*.ncu-rep
