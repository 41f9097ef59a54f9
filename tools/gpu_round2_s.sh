#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu_s.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_s.log
KEYS=conv1_fwd,conv2_dgrad,conv1_wgrad bash tools/gpu_ab.sh
for v in old new; do
PAACB_LIB=$PWD/build/ab/libpaacb_$v.so timeout 300 python bench.py --arch NIPS --steps 20 --no_cpu_baseline --no_variants --no_e2e > gpurun_out/abn_$v.json 2> gpurun_out/ab.err
python - <<PY
import json
d=json.loads(open('gpurun_out/abn_$v.json').read().strip().splitlines()[-1])
ks={x['name']:x['ms']/d['steps'] for x in d['kernels']}
print('NIPS $v ms/step %.3f'%d['ms_per_step'], ' '.join('%s %.3f'%(k,ks[k]) for k in ('conv1_fwd','conv2_dgrad')), 'clocks', d['clocks']['sm_mhz'])
PY
done
