for nt in 0 1; do for th in 12 15; do export TL_E2E_NT=$nt
TL_E2E_THREADS=$th python bench.py --steps 20 --warmup 5 --no-extra --no-cpu > gpurun_out/r2c_bench_${nt}_$th.json 2> gpurun_out/r2c_bench_${nt}_$th.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2c_bench_${nt}_$th.json').read().strip().splitlines()[-1])
e=d['e2e']; print('nt $nt threads $th', round(e['value']/1e6,1), e['api'][-40:])
for k,v in e['formats'].items(): print('  ',k, round(v['value']/1e6,1), round(v['ms_per_step'],4), v['host_ms_per_step'])
"
done; done
nproc; lscpu | grep -i "model name\|^CPU(s)\|Thread\|Socket\|L3"
