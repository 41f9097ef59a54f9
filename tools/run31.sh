timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_update.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest_gpu.log
bash tools/run_ab.sh
