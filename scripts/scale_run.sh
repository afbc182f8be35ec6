set -u
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_scale_$N.json 2> gpurun_out/r2_scale_$N.err
tail -c 300 gpurun_out/r2_scale_$N.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_scale_$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], {k:v.get('value') for k,v in d.get('workloads',{}).items()})
"
