#!/bin/bash
# round 2, pass T: K1 copy pipeline -- parity tests, then same-box A/B by environment knob (PAACB_K1_PIPE, PAACB_K1_HINTS)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_engine.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu_t.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_t.log
for rep in 1 2; do for cfg in "0 0" "1 0" "1 3" "1 1" "1 2"; do
set -- $cfg
PAACB_K1_PIPE=$1 PAACB_K1_HINTS=$2 timeout 300 python bench.py --steps 20 --no_cpu_baseline --no_variants --no_e2e > gpurun_out/k1ab_$1_$2.json 2> gpurun_out/k1ab.err || tail -5 gpurun_out/k1ab.err
python - <<PY
import json
d=json.loads(open('gpurun_out/k1ab_$1_$2.json').read().strip().splitlines()[-1])
ks={x['name']:x['ms']/d['steps'] for x in d['kernels']}
print('pipe=$1 hints=$2 ms/step %.3f'%d['ms_per_step'], ' '.join('%s %.3f'%(k,ks[k]) for k in ('preprocess_u8','conv1_fwd','conv2_fwd')), 'clocks', d['clocks']['sm_mhz'])
PY
done; done
for cfg in "0 0" "1 3"; do
set -- $cfg
PAACB_K1_PIPE=$1 PAACB_K1_HINTS=$2 timeout 300 python tools/microbench_cfg5.py > gpurun_out/micro_k1_$1_$2.json 2> gpurun_out/micro.err || tail -5 gpurun_out/micro.err
python - <<PY
import json
d=json.load(open('gpurun_out/micro_k1_$1_$2.json'))
rows=d['rows'] if isinstance(d,dict) and 'rows' in d else d
for r in rows:
    if 'K1' in r['kernel'] or 'preprocess' in r['kernel']: print('micro pipe=$1', r['kernel'][:40], '%.1f us'%r['median_us'], 'moved_frac %.3f'%r['moved_frac'])
PY
done
