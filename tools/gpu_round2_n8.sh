#!/bin/bash
# round-2 GPU pass on 8 B200s: G-rank update == 1-rank update (kept logs), torchrun train() ranks identical, concurrent
# host-link probe at 4 and 8 ranks, N = 8 bench line (+ the variant without the overlapped all-reduce)
G=8
mkdir -p gpurun_out
OUT=gpurun_out/multi_gpu_check_w${G}.log
: > $OUT
for cfg in "fp32 NATURE" "tf32x3 NATURE" "bf16x3 NATURE" "bf16x3 NIPS"; do
  set -- $cfg
  PAACB_CHECK_MATH=$1 PAACB_CHECK_ARCH=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 \
    --master-port 29533 tools/multi_gpu_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tee -a $OUT | tail -2
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29547 \
  tools/train_ranks_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tee -a $OUT | tail -2
for g in 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29561 \
    tools/experiments/pcie_concurrent.py 2>/dev/null | tail -1 > gpurun_out/pcie_concurrent_w$g.json; echo "pcie w$g rc=$?"; cat gpurun_out/pcie_concurrent_w$g.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus $G --steps 30 > gpurun_out/r02_bench_bf16x3_n$G.json 2> gpurun_out/bench_n$G.err; echo "bench rc=$?"
tail -2 gpurun_out/bench_n$G.err
python tools/show_bench.py gpurun_out/r02_bench_bf16x3_n$G.json 2>&1 | head -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29565 bench.py --gpus $G --steps 30 --no_e2e --no_variants --no_overlap_allreduce > gpurun_out/r02_bench_bf16x3_n${G}_no_overlap.json 2> gpurun_out/bench_n${G}b.err; echo "bench (no overlap) rc=$?"
python tools/show_bench.py gpurun_out/r02_bench_bf16x3_n${G}_no_overlap.json 2>&1 | head -2
PAACB_SM_RESERVE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29567 bench.py --gpus $G --steps 30 --no_e2e --no_variants > gpurun_out/r02_bench_bf16x3_n${G}_reserve0.json 2> gpurun_out/bench_n${G}c.err; echo "bench (reserve 0) rc=$?"
python tools/show_bench.py gpurun_out/r02_bench_bf16x3_n${G}_reserve0.json 2>&1 | head -2
