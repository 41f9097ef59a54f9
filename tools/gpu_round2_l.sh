#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu_l.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_l.log
for a in NATURE NIPS; do
timeout 600 python bench.py --arch $a --steps 30 --no_e2e --no_cpu_baseline --no_variants > gpurun_out/bench_l_$a.json 2> gpurun_out/bench_l.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/bench_l_$a.json 2>&1 | grep -E "value|heads|clocks"
done
