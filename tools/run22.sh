for d in 0 64 128 192 224; do
PAACB_DBG=$d timeout 300 python bench.py --steps 10 --no_cpu_baseline --no_variants --no_e2e > gpurun_out/abl_$d.json 2> gpurun_out/abl.err
python - <<PY
import json
d=json.loads(open('gpurun_out/abl_$d.json').read().strip().splitlines()[-1])
ks={x['name']:x['ms']/d['steps'] for x in d['kernels']}
print('dbg=$d', ' '.join('%s %.3f'%(k,ks[k]) for k in ('conv1_wgrad','conv2_wgrad','conv3_wgrad')), 'clocks', d['clocks']['sm_mhz'])
PY
done
