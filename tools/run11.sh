timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_update.py -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/pytest_gpu2.log
for flag in "" "--no_overlap_allreduce"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --no_e2e --no_variants $flag > gpurun_out/bench_n2$flag.json 2> gpurun_out/bench_n2.err; echo bench rc=$?
tail -3 gpurun_out/bench_n2.err
python tools/show_bench.py gpurun_out/bench_n2$flag.json 2>&1 | head -3
done
