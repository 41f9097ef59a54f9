timeout 600 python bench.py --no_cpu_baseline > gpurun_out/bench_bf16x3.log 2> gpurun_out/bench_bf16x3.err; echo bench rc=$?
tail -2 gpurun_out/bench_bf16x3.err
