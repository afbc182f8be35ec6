#!/bin/bash
# Collect the ncu evidence for profiles/ on the GPU box (run under gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash scripts/collect_profiles.sh rNN'
# 1. launch list (gpu__time_duration.sum) of the default bench command
# 2. one --set full capture of the tick kernel per workload (after the same command passed without ncu)
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 64 --warmup 32 --no-cpu --e2e-steps 1"
$BENCH > $OUT/${TAG}_plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_c2.csv $BENCH > $OUT/${TAG}_ncu_launches.log 2>&1
for W in c2 c3-compact c3-full c4; do
  CMD="python bench.py --steps 8 --warmup 4 --ticks-per-launch 4 --workload $W --no-cpu --e2e-steps 1"
  [ "$W" = "c2" ] && CMD="python bench.py --steps 64 --warmup 32 --ticks-per-launch 32 --workload c2 --no-cpu --e2e-steps 1"
  $CMD > $OUT/${TAG}_plain_$W.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:tick_kernel -s 1 -c 1 -o $OUT/${TAG}_prof_$W $CMD > $OUT/${TAG}_ncu_$W.log 2>&1
  tail -1 $OUT/${TAG}_ncu_$W.log
done
ls -la $OUT | tail -20
