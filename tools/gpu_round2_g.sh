#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipe.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu_pipe.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu_pipe.log
timeout 300 python tools/experiments/pipe_tune.py > gpurun_out/pipe_tune.json 2> gpurun_out/pipe_tune.err; echo "tune rc=$?"; grep arch gpurun_out/pipe_tune.err
