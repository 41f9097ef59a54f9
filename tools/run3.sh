timeout 300 python tools/bf16_check.py fwd > gpurun_out/bf16_fwd.log 2>&1; echo fwd rc=$?
timeout 300 python tools/bf16_check.py bwd > gpurun_out/bf16_bwd.log 2>&1; echo bwd rc=$?
grep -E "^FWD|^BWD" gpurun_out/bf16_fwd.log gpurun_out/bf16_bwd.log
timeout 600 python bench.py --no_cpu_baseline > gpurun_out/bench_bf16x3.log 2> gpurun_out/bench_bf16x3.err; echo bench rc=$?
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q > gpurun_out/pytest_tc.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest_tc.log
