#!/bin/bash
# Same-box comparison of several builds of the library (box-to-box variation is ~3 %, larger than most kernel tweaks):
#   gpurun -- 'bash scripts/ab.sh "<libA.so> <libB.so> ..." c2 c3-compact ...'   (3 interleaved runs per workload;
#   the first library is the base of the printed ratios)
LIBS=($1); shift
for W in "$@"; do
  declare -a SUM; for i in "${!LIBS[@]}"; do SUM[$i]=0; done
  for R in 1 2 3; do
    for i in "${!LIBS[@]}"; do
      V=$(TOWNLET_B200_LIB=${LIBS[$i]} python bench.py --workload $W --no-cpu --no-e2e --no-extra 2>/dev/null | tail -1 | python -c "import sys,json; print(json.loads(sys.stdin.read())['value'])")
      SUM[$i]=$(python -c "print(${SUM[$i]} + $V)")
    done
  done
  for i in "${!LIBS[@]}"; do
    python -c "a=${SUM[0]}/3e9; b=${SUM[$i]}/3e9; print(f'$W  $(basename ${LIBS[$i]})  {b:.3f} G  vs base {b/a:.4f}')"
  done
done
