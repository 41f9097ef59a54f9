#!/bin/bash
# Multi-GPU equivalence suite; run with `gpurun --gpus G -- bash tools/gpu_multi_suite.sh G` (G = 2 or 8).  Keeps the stdout of
# every check under gpurun_out/ (copied to profiles/r02_multi_gpu_check_w<G>.log): G-rank update == 1-rank update on the
# concatenated batch in all three arithmetic modes, and PAACLearner.train() under torchrun ends with identical ranks.
G=${1:-2}
OUT=gpurun_out/multi_gpu_check_w${G}.log
: > $OUT
for m in fp32 tf32x3 bf16x3; do
  PAACB_CHECK_MATH=$m timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 \
    --master-port 29533 tools/multi_gpu_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tee -a $OUT | tail -3
done
for a in NATURE NIPS; do
  PAACB_CHECK_MATH=bf16x3 PAACB_CHECK_ARCH=$a timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G \
    --master-addr 127.0.0.1 --master-port 29535 tools/multi_gpu_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tee -a $OUT | tail -2
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29547 \
  tools/train_ranks_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tee -a $OUT | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29549 \
  tools/train_ranks_check.py --graphs false --train_forward batched 2>&1 | grep -v "^W\|^\*\*\*" | tee -a $OUT | tail -3
grep -c "ok$" $OUT
