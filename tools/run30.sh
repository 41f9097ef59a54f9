PAACB_DBG=32768 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu.log
for d in 32768 0; do
PAACB_DBG=$d timeout 300 python bench.py --steps 20 --no_cpu_baseline --no_variants --no_e2e > gpurun_out/abl_$d.json 2> gpurun_out/abl.err
python - <<PY
import json
d=json.loads(open('gpurun_out/abl_$d.json').read().strip().splitlines()[-1])
ks={x['name']:x['ms']/d['steps'] for x in d['kernels']}
print('dbg=$d ms/step %.3f'%d['ms_per_step'], ' '.join('%s %.3f'%(k,ks[k]) for k in ('conv1_fwd','conv2_fwd')), 'clocks', d['clocks']['sm_mhz'])
PY
done
