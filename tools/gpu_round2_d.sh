#!/bin/bash
# round-2 GPU pass D: weight-gradient K-step skipping / partial N-concatenation (parity + timing), fc no-load ablation
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_update.py tests/test_gpu_fullsize.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_d.log
timeout 600 python bench.py --steps 20 --no_e2e --no_cpu_baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_d.err
python tools/show_bench.py gpurun_out/bench_d.json 2>&1 | tail -22
timeout 600 python bench.py --arch NIPS --steps 20 --no_e2e --no_cpu_baseline > gpurun_out/bench_d_nips.json 2> gpurun_out/bench_d_nips.err; echo "bench nips rc=$?"
python tools/show_bench.py gpurun_out/bench_d_nips.json 2>&1 | tail -18
timeout 600 python tools/ablation_sweep.py 2048 > gpurun_out/r02_ablations_fc.json 2> gpurun_out/abl.err; echo "ablation rc=$?"; tail -3 gpurun_out/abl.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_ablations_fc.json'))
for r in d['rows']: print(r['PAACB_DBG'], r['ms_per_step'])
PY
