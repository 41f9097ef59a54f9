#!/bin/bash
# round 2, pass U: device-resident arm, whole batch on one stream vs environment slices on their own streams (same box)
mkdir -p gpurun_out
for rep in 1 2; do for S in 1 2 4; do
timeout 300 python bench.py --steps 20 --dev_slices $S --no_cpu_baseline --no_variants --no_e2e > gpurun_out/slices_$S.json 2> gpurun_out/slices.err || tail -5 gpurun_out/slices.err
python - <<PY
import json
d=json.loads(open('gpurun_out/slices_$S.json').read().strip().splitlines()[-1])
ks={x['name']:x['ms']/d['steps'] for x in d['kernels']}
print('slices=$S ms/step %.3f value %.0f launches %d profiled %.3f'%(d['ms_per_step'], d['value'], d['gpu_launches'], d['profiled_ms_per_step']), 'loss %.6f'%d['loss'], 'clocks', d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
done; done
for S in 1 2; do
timeout 300 python bench.py --arch NIPS --steps 20 --dev_slices $S --no_cpu_baseline --no_variants --no_e2e > gpurun_out/slices_nips_$S.json 2> gpurun_out/slices.err || tail -5 gpurun_out/slices.err
python - <<PY
import json
d=json.loads(open('gpurun_out/slices_nips_$S.json').read().strip().splitlines()[-1])
print('NIPS slices=$S ms/step %.3f value %.0f'%(d['ms_per_step'], d['value']), 'loss %.6f'%d['loss'])
PY
done
