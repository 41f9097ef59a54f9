#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
  for pdl in 1 0; do
    PAACB_PDL=$pdl timeout 600 python bench.py --steps 30 --no_e2e --no_cpu_baseline --no_variants > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err
    python - <<PY
import json
d=json.loads(open('gpurun_out/bench_k.json').read().strip().split('\n')[-1])
print('NATURE pdl=$pdl run $i: %.4f ms  %d MHz' % (d['ms_per_step'], d['clocks']['sm_mhz']))
PY
  done
done
for pdl in 1 0; do
  PAACB_PDL=$pdl timeout 300 python tools/small_batch.py --no_learner --out gpurun_out/small_batch_pdl$pdl.json > /dev/null 2>&1; python - <<PY
import json
for e in json.load(open('gpurun_out/small_batch_pdl$pdl.json'))['engine']: print('pdl=$pdl', e['arch'], e['graphs'], e['train_forward'], round(e['ms_per_cycle'],4), round(e['update_ms'],4))
PY
done
