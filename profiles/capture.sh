#!/bin/bash
# profiles/capture.sh <tag> <config> [kernel-regex ...] -- run under gpurun on ONE GPU.
# 1. launch list of a short bench run (cold-cache, serialised: compare SHARES, not absolutes)
# 2. one `ncu --set full` capture per named kernel (the 3rd launch of each)
# Outputs land in gpurun_out/ (scratch); summaries are copied into profiles/ by hand.
set -u
TAG=${1:-r01}; CFG=${2:-c2}; shift 2 || true
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
    --log-file gpurun_out/launches_${TAG}_${CFG}.csv python bench.py --config ${CFG} --steps 1 --warmup 1 --profile-only \
    > gpurun_out/bench_under_ncu_${TAG}_${CFG}.log 2>&1
for K in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:${K} -s 2 -c 1 \
      -o gpurun_out/prof_${TAG}_${CFG}_${K} -f python bench.py --config ${CFG} --steps 1 --warmup 0 --profile-only \
      > gpurun_out/ncu_${TAG}_${CFG}_${K}.log 2>&1
done
ls -la gpurun_out | tail -20
