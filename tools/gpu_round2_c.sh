#!/bin/bash
# round-2 GPU pass C (2 GPUs): multi-GPU equivalence suite with kept logs, the -m gpu tests that need 2 devices,
# concurrent host-link probe, N=2 bench line
G=${1:-2}
mkdir -p gpurun_out
bash tools/gpu_multi_suite.sh $G
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_engine.py -m gpu -q -rA -p no:cacheprovider > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest rc=$?"
grep -E "^(PASSED|FAILED|ERROR|SKIPPED)|passed|failed" gpurun_out/pytest_gpu_c.log | tail -30
for g in 1 $G; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29561 \
    tools/experiments/pcie_concurrent.py 2>/dev/null | tail -1 > gpurun_out/pcie_concurrent_w$g.json; echo "pcie w$g rc=$?"; cat gpurun_out/pcie_concurrent_w$g.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus $G --steps 20 > gpurun_out/r02_bench_bf16x3_n$G.json 2> gpurun_out/bench_n$G.err; echo "bench rc=$?"
tail -2 gpurun_out/bench_n$G.err
python tools/show_bench.py gpurun_out/r02_bench_bf16x3_n$G.json 2>&1 | head -3
