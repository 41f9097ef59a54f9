timeout 300 python tools/microbench_cfg5.py > gpurun_out/micro_cfg5.json 2> gpurun_out/micro_cfg5.err; echo micro rc=$?; tail -3 gpurun_out/micro_cfg5.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/micro_cfg5.json'))
for r in d['rows']: print('%-70s med %.1f best %.1f us  contract %.0f GB/s (%.2f)  moved %.0f GB/s (%.2f)'%(r['kernel'][:70], r['median_us'], r['best_us'], r['achieved_gbs'], r['frac'], r['moved_gbs'], r['moved_frac']))
PY
PAACB_OPT_TWO_PASS=1 timeout 300 python tools/microbench_cfg5.py --iters 10 2>&1 | grep -A3 '"kernel": "clip' | grep "kernel\|median"
