#!/bin/bash
# round 2, final build on one 8-GPU box: N = 1 (quick, same box) and the N = 8 bench line
mkdir -p gpurun_out
timeout 300 python bench.py --steps 30 --no_cpu_baseline --no_variants --no_e2e > gpurun_out/n8box_n1.json 2> gpurun_out/n8box.err; echo "n1 rc=$?"
python tools/show_bench.py gpurun_out/n8box_n1.json 2>&1 | head -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 8 --steps 30 > gpurun_out/r02_bench_bf16x3_n8.json 2> gpurun_out/bench_n8.err; echo "bench rc=$?"
tail -2 gpurun_out/bench_n8.err
python tools/show_bench.py gpurun_out/r02_bench_bf16x3_n8.json 2>&1 | head -8
