for d in 0 1 2 4 8 3 12 15; do
PAACB_DBG=$d timeout 300 python bench.py --steps 10 --no_cpu_baseline --no_variants --no_e2e > gpurun_out/abl_$d.json 2> gpurun_out/abl.err
python - <<PY
import json
d=json.loads(open('gpurun_out/abl_$d.json').read().strip().splitlines()[-1])
k=[x for x in d['kernels'] if x['name']=='conv1_fwd'][0]
print('dbg=$d conv1_fwd ms/step %.3f  clocks %s'%(k['ms']/d['steps'], d['clocks']['sm_mhz']))
PY
done
