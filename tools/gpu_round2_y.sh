#!/bin/bash
# round 2, pass Y: end-to-end arm (pinned host frames), zero-copy K1 as CTA-per-env kernel (grid 96) vs through the TMA pipeline
mkdir -p gpurun_out
for rep in 1 2; do for cfg in "1 16" "2 8" "2 16" "2 32"; do
set -- $cfg
PAACB_K1_PIPE=$1 PAACB_K1_PIPE_HOST_GRID=$2 timeout 300 python bench.py --steps 20 --no_cpu_baseline --no_variants > gpurun_out/e2e_$1_$2.json 2> gpurun_out/e2e.err || tail -5 gpurun_out/e2e.err
python - <<PY
import json
d=json.loads(open('gpurun_out/e2e_$1_$2.json').read().strip().splitlines()[-1])
print('k1_pipe=$1 grid=$2 ms/step %.3f  e2e %.0f env-steps/s %.3f ms'%(d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']), 'loss %.6f'%d['loss'])
PY
done; done
