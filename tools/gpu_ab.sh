# same-box A/B of two builds of the library: build/ab/libpaacb_old.so vs build/ab/libpaacb_new.so (see PAACB_LIB in paac_b200/_lib.py)
KEYS=${KEYS:-conv1_fwd,conv2_dgrad,conv3_dgrad,fc4_dgrad}
for rep in 1 2; do for v in old new; do
PAACB_LIB=$PWD/build/ab/libpaacb_$v.so timeout 300 python bench.py --steps 20 --no_cpu_baseline --no_variants --no_e2e > gpurun_out/ab_$v.json 2> gpurun_out/ab.err
KEYS=$KEYS python - <<PY
import json, os
d=json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1])
ks={x['name']:x['ms']/d['steps'] for x in d['kernels']}
print('$v ms/step %.3f'%d['ms_per_step'], ' '.join('%s %.3f'%(k,ks[k]) for k in os.environ['KEYS'].split(',')), 'clocks', d['clocks']['sm_mhz'])
PY
done; done
