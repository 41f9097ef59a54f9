#!/bin/bash
# round-2 GPU pass I: programmatic dependent launch -- whole GPU suite (results must not change), bench Nature + NIPS, A/B vs PAACB_PDL=0
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu_i.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_i.log
for pdl in 1 0; do
  PAACB_PDL=$pdl timeout 600 python bench.py --steps 20 --no_e2e --no_cpu_baseline --no_variants > gpurun_out/bench_i_pdl$pdl.json 2> gpurun_out/bench_i.err; echo "bench pdl=$pdl rc=$?"
  python tools/show_bench.py gpurun_out/bench_i_pdl$pdl.json 2>&1 | head -1
  PAACB_PDL=$pdl timeout 600 python bench.py --arch NIPS --steps 20 --no_e2e --no_cpu_baseline --no_variants > gpurun_out/bench_i_nips_pdl$pdl.json 2> gpurun_out/bench_i.err; echo "bench nips pdl=$pdl rc=$?"
  python tools/show_bench.py gpurun_out/bench_i_nips_pdl$pdl.json 2>&1 | head -1
done
PAACB_PDL=1 timeout 600 python bench.py --steps 20 --no_e2e --no_cpu_baseline --no_variants > gpurun_out/bench_i_pdl1b.json 2> gpurun_out/bench_i.err; python tools/show_bench.py gpurun_out/bench_i_pdl1b.json 2>&1 | head -1
timeout 300 python tools/small_batch.py --no_learner --out gpurun_out/small_batch_pdl.json > /dev/null 2>&1; python - <<'PY'
import json
for e in json.load(open('gpurun_out/small_batch_pdl.json'))['engine']: print(e['arch'], e['graphs'], e['train_forward'], round(e['ms_per_cycle'],4), round(e['update_ms'],4))
PY
