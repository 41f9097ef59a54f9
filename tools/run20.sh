timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu.log
for v in 1 0 1 0; do
PAACB_NO_PDL=$v timeout 300 python bench.py --steps 40 --no_cpu_baseline --no_variants --no_e2e > gpurun_out/pdl_$v.json 2> gpurun_out/pdl.err
python - <<PY
import json
d=json.loads(open('gpurun_out/pdl_$v.json').read().strip().splitlines()[-1])
print('NO_PDL=$v ms/step %.4f value %.0f profiled %.4f clocks %s'%(d['ms_per_step'], d['value'], d['profiled_ms_per_step'], d['clocks']['sm_mhz']))
PY
done
