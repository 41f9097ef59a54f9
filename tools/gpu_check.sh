timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench rc=$?
tail -3 gpurun_out/bench_n1.err
python tools/show_bench.py gpurun_out/bench_n1.json 2>&1 | tail -24
