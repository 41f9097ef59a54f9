nvidia-smi topo -m 2>/dev/null | head -8
timeout 600 python bench.py --no_cpu_baseline > gpurun_out/bench_numa_n1.json 2> gpurun_out/bench_n1.err; echo bench1 rc=$?
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --no_cpu_baseline > gpurun_out/bench_numa_n2.json 2> gpurun_out/bench_n2.err; echo bench2 rc=$?
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu.log
