bash scripts/tpc_probe.sh c3-compact "0 1 2 3"
for P in 0 16 24 32; do
  V=$(python bench.py --workload c4 --no-cpu --no-e2e --no-extra --option pass_agents=$P 2>/dev/null | tail -1 | python -c "import sys,json; print(json.loads(sys.stdin.read())['value'])")
  echo "c4 pass_agents=$P: $V"
done
for P in 16 20; do
  V=$(python bench.py --workload c3-compact --no-cpu --no-e2e --no-extra --option pass_agents=$P 2>/dev/null | tail -1 | python -c "import sys,json; print(json.loads(sys.stdin.read())['value'])")
  echo "c3-compact pass_agents=$P: $V"
done
