nvidia-smi -L | wc -l
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo bench rc=$?
tail -2 gpurun_out/bench_n$N.err
python tools/show_bench.py gpurun_out/bench_n$N.json 2>&1 | head -3
