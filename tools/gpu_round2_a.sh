#!/bin/bash
# round-2 GPU pass A: tcgen05 peaks, the whole -m gpu suite (all failures, with the parity lines), smoke, one short bench
mkdir -p gpurun_out
timeout 120 tools/probe/umma_peak > gpurun_out/umma_peaks.json 2> gpurun_out/umma_peaks.err; echo "umma_peak rc=$?"; cat gpurun_out/umma_peaks.json
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/pytest_gpu_a.log 2>&1; echo "pytest rc=$?"
grep -E "^(PASSED|FAILED|ERROR|SKIPPED)|passed|failed" gpurun_out/pytest_gpu_a.log | tail -60
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_a.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_a.log
timeout 600 python bench.py --steps 20 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_a.err
python tools/show_bench.py gpurun_out/bench_a.json 2>&1 | tail -24
