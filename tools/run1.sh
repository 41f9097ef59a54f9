set -x
nvidia-smi -L
timeout 900 python tools/probe/run_probe.py > gpurun_out/probe.log 2>&1; echo probe rc=$?
timeout 600 python bench.py > gpurun_out/r01_bench_default.log 2> gpurun_out/r01_bench_default.err; echo bench rc=$?
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_reference.log 2>&1; echo ref rc=$?
python bench.py --steps 2 --warmup 3 --no_cpu_baseline --no_e2e > gpurun_out/plain_ncu.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r01_tc_launches.csv python bench.py --steps 2 --warmup 3 --no_cpu_baseline --no_e2e > gpurun_out/ncu_launch.log 2>&1; echo ncu rc=$?
