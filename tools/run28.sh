timeout 900 python -m pytest tests/test_gpu_learner.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_learner.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/pytest_learner.log
