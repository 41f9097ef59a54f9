#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu_w.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_w.log
