for d in 0 256 512 1024 768; do
PAACB_DBG=$d timeout 300 python bench.py --steps 10 --no_cpu_baseline --no_variants --no_e2e > gpurun_out/abl_$d.json 2> gpurun_out/abl.err
python - <<PY
import json
d=json.loads(open('gpurun_out/abl_$d.json').read().strip().splitlines()[-1])
ks={x['name']:x['ms']/d['steps'] for x in d['kernels']}
print('dbg=$d', ' '.join('%s %.3f'%(k,ks[k]) for k in ('conv2_fwd','conv3_fwd','conv3_dgrad','conv2_dgrad')), 'clocks', d['clocks']['sm_mhz'])
PY
done
