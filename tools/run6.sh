nvidia-smi -L | wc -l
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --no_cpu_baseline > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo bench4 rc=$?; tail -2 gpurun_out/bench_n4.err
