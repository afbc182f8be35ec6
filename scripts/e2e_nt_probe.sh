#!/bin/bash
# e2e (host-buffer path) at N ranks with and without non-temporal stores in the host expander:
#   gpurun --gpus 8 -- 'bash scripts/e2e_nt_probe.sh 8'
N=$1
for NT in 1 0; do
  TL_E2E_NT=$NT timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 --no-extra --no-cpu > gpurun_out/e2e_nt${NT}_$N.json 2> gpurun_out/e2e_nt${NT}_$N.err
  python -c "
import json
d=json.loads(open('gpurun_out/e2e_nt${NT}_$N.json').read().strip().splitlines()[-1])
e=d['e2e']; print('NT=$NT N=$N e2e', round(e['value']/1e6,1), {k[:26]:round(v['value']/1e6,1) for k,v in e['formats'].items()})
"
done
