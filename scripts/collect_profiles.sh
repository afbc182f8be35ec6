#!/bin/bash
# Collect the ncu evidence for profiles/ on the GPU box (run under gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash scripts/collect_profiles.sh rNN'
# 1. launch list (gpu__time_duration.sum) of the driver's bench command (shortened: fewer repetitions, no CPU / e2e legs)
# 2. one --set full capture of the tick kernel per workload and one of the fused policy kernel
#    (each after the same command passed without ncu)
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
SHORT="--no-cpu --no-extra --no-e2e --min-launches 40 --min-seconds 0.01"
BENCH="python bench.py --steps 20 --warmup 5 $SHORT"
$BENCH > $OUT/${TAG}_plain_c2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_c2.csv $BENCH > $OUT/${TAG}_ncu_launches.log 2>&1
C5="python bench.py --workload c5 --steps 20 --warmup 5 --ticks-per-launch 8 --no-cpu --no-extra --no-e2e --min-seconds 0.005"
$C5 > $OUT/${TAG}_plain_c5.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_c5.csv $C5 > $OUT/${TAG}_ncu_launches_c5.log 2>&1
for W in c2 c3-compact c3-full c4; do
  TPL=4; [ "$W" = "c2" ] && TPL=32
  CMD="python bench.py --steps 20 --warmup 4 --ticks-per-launch $TPL --workload $W --no-cpu --no-extra --no-e2e --min-launches 4 --min-seconds 0.001"
  $CMD > $OUT/${TAG}_plain_$W.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:tick_kernel -s 2 -c 1 -o $OUT/${TAG}_prof_$W $CMD > $OUT/${TAG}_ncu_$W.log 2>&1
  tail -1 $OUT/${TAG}_ncu_$W.log
done
POL="python -m pytest tests/test_policy_step_gpu.py -m gpu -q -k 65613"
$POL > $OUT/${TAG}_plain_policy.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:policy_mlp -c 1 -o $OUT/${TAG}_prof_policy $POL > $OUT/${TAG}_ncu_policy.log 2>&1
tail -1 $OUT/${TAG}_ncu_policy.log
ls -la $OUT | tail -12
