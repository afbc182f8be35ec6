#!/bin/bash
# Same-box A/B of two builds of the library (box-to-box variation is ~3 %, larger than most kernel tweaks):
#   gpurun -- 'bash scripts/ab.sh <libA.so> <libB.so> c2 c3-compact ...'   (3 alternating runs per workload)
A=$1; B=$2; shift 2
for W in "$@"; do
  for L in A B; do eval "SUM_$L=0"; done
  for R in 1 2 3; do
    for L in A B; do
      LIB=$A; [ $L = B ] && LIB=$B
      V=$(TOWNLET_B200_LIB=$LIB python bench.py --workload $W --no-cpu --no-e2e --no-extra 2>/dev/null | tail -1 | python -c "import sys,json; print(json.loads(sys.stdin.read())['value'])")
      eval "SUM_$L=\$(python -c \"print(\$SUM_$L + $V)\")"
    done
  done
  python -c "a=$SUM_A/3e9; b=$SUM_B/3e9; print(f'$W  A {a:.3f} G  B {b:.3f} G  B/A {b/a:.4f}')"
done
