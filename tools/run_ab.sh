# same-box A/B of two builds of the library: build/ab/libpaacb_old.so vs build/ab/libpaacb_new.so
for rep in 1; do for v in old new4 old new4; do
PAACB_LIB=$PWD/build/ab/libpaacb_$v.so timeout 300 python bench.py --steps 20 --no_cpu_baseline --no_variants --no_e2e > gpurun_out/ab_$v.json 2> gpurun_out/ab.err
python - <<PY
import json
d=json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1])
ks={x['name']:x['ms']/d['steps'] for x in d['kernels']}
print('$v ms/step %.3f'%d['ms_per_step'], ' '.join('%s %.3f'%(k,ks[k]) for k in ('conv1_fwd','conv2_dgrad','conv2_fwd')), 'clocks', d['clocks']['sm_mhz'])
PY
done; done
