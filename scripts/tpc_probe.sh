#!/bin/bash
# towns per CTA of the role kernel on a workload (default: chosen from the batch shape):  bash scripts/tpc_probe.sh c5 "0 2 3 4"
W=$1
for T in $2; do
  V=$(python bench.py --workload $W --no-cpu --no-e2e --no-extra --option towns_per_cta=$T 2>/dev/null | tail -1 | python -c "import sys,json; print(json.loads(sys.stdin.read())['value'])")
  echo "$W towns_per_cta=$T: $V"
done
