#!/bin/bash
# round 2, pass Z: forward conv kernels, epilogue-issued L2 prefetch of the tile after next (PAACB_FWD_PREFETCH bit 0 conv2, bit 1 conv3)
mkdir -p gpurun_out
for rep in 1 2; do for K in 0 1 2 3; do
PAACB_FWD_PREFETCH=$K timeout 300 python bench.py --steps 20 --no_cpu_baseline --no_variants --no_e2e > gpurun_out/fpf_$K.json 2> gpurun_out/fpf.err || tail -5 gpurun_out/fpf.err
python - <<PY
import json
d=json.loads(open('gpurun_out/fpf_$K.json').read().strip().splitlines()[-1])
ks={x['name']:x['ms']/d['steps'] for x in d['kernels']}
print('fwd_prefetch=$K ms/step %.3f'%d['ms_per_step'], ' '.join('%s %.3f'%(k,ks[k]) for k in ('conv2_fwd','conv3_fwd','conv1_fwd','fc4_fwd')), 'loss %.6f'%d['loss'], 'clocks', d['clocks']['sm_mhz'])
PY
done; done
for K in 0 3; do
PAACB_FWD_PREFETCH=$K timeout 300 python bench.py --arch NIPS --steps 20 --no_cpu_baseline --no_variants --no_e2e > gpurun_out/fpfn_$K.json 2> gpurun_out/fpf.err || tail -5 gpurun_out/fpf.err
python - <<PY
import json
d=json.loads(open('gpurun_out/fpfn_$K.json').read().strip().splitlines()[-1])
ks={x['name']:x['ms']/d['steps'] for x in d['kernels']}
print('NIPS fwd_prefetch=$K ms/step %.3f'%d['ms_per_step'], ' '.join('%s %.3f'%(k,ks[k]) for k in ('conv2_fwd','conv1_fwd')), 'loss %.6f'%d['loss'])
PY
done
