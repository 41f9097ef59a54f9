nvidia-smi -L
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo smoke rc=$?; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/r01_bench_bf16x3_n1.json 2> gpurun_out/bench_n1.err; echo bench1 rc=$?
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > gpurun_out/r01_bench_bf16x3_n2.json 2> gpurun_out/bench_n2.err; echo bench2 rc=$?
python bench.py --steps 2 --warmup 3 --no_cpu_baseline --no_e2e > gpurun_out/plain_ncu.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r01_bf16x3_launches.csv python bench.py --steps 2 --warmup 3 --no_cpu_baseline --no_e2e > gpurun_out/ncu_launch.log 2>&1; echo ncu rc=$?
