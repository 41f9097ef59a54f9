timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/pytest_full.log 2>&1; echo pytest rc=$?; tail -15 gpurun_out/pytest_full.log
