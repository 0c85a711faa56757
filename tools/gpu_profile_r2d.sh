#!/bin/bash
# Round-2 final evidence run on the GPU box: tests -> full bench line -> ncu launch list of two timed steps -> ncu --set full of the FlowLM
# attention kernels, the fused flow head and the fused SEANet tail. $1 = tag.
TAG=${1:-r2d}
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
ARGS="--steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
python bench.py $ARGS > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
PTTS_NCU_RANGE=1 timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py $ARGS > gpurun_out/ncu_$TAG.log 2>&1
PTTS_NCU_RANGE=1 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"attn_(tile|flow_split|merge)" -c 6 -f -o gpurun_out/full_attn_$TAG \
  python bench.py $ARGS > gpurun_out/ncu_full_attn_$TAG.log 2>&1
PTTS_NCU_RANGE=1 timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:"head_res_cluster|seanet_tail|pcm_combine" -c 3 -f -o gpurun_out/full_fused_$TAG \
  python bench.py $ARGS > gpurun_out/ncu_full_fused_$TAG.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
