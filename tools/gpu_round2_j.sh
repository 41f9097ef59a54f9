#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_j.log 2>&1; echo "pytest (pdl default) rc=$?"
tail -6 gpurun_out/pytest_gpu_j.log
for pdl in 0 1; do
  for i in 1 2 3 4; do
    PAACB_PDL=$pdl timeout 300 python -m pytest tests/test_gpu_engine.py -k graph_replay -m gpu -q -p no:cacheprovider 2>&1 | grep -E "scaled max error|passed|failed" | tr '\n' ' '; echo " [pdl=$pdl run $i]"
  done
done
