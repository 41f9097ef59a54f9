timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/microbench_cfg5.py > gpurun_out/micro_cfg5.json 2> gpurun_out/micro_cfg5.err; echo micro rc=$?; tail -3 gpurun_out/micro_cfg5.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/micro_cfg5.json'))
for r in d['rows']: print('%-70s med %.1f us  contract %.0f GB/s (%.2f)  moved %.0f GB/s (%.2f)'%(r['kernel'][:70], r['median_us'], r['achieved_gbs'], r['frac'], r['moved_gbs'], r['moved_frac']))
PY
PAACB_OPT_TWO_PASS=1 timeout 300 python tools/microbench_cfg5.py --iters 10 2>&1 | grep -A3 '"kernel": "clip' | grep "kernel\|median"
timeout 600 python bench.py --no_cpu_baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench rc=$?
tail -3 gpurun_out/bench_n1.err
python tools/show_bench.py gpurun_out/bench_n1.json 2>&1 | tail -24
