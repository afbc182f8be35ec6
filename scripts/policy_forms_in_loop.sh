for W in c5 c5-bf16; do
  for O in "policy_pair=0" "policy_pair=0 --option policy_multicast=1" "policy_pair=1"; do
    V=$(python bench.py --workload $W --no-cpu --no-e2e --no-extra --option $O 2>/dev/null | tail -1 | python -c "import sys,json; print(json.loads(sys.stdin.read())['value'])")
    echo "$W $O: $V"
  done
done
