#!/bin/bash
# round-2 evidence pass (1 GPU): whole -m gpu suite, smoke, full bench lines (Nature, NIPS, reference arm), ncu launch list and
# one ncu --set full capture of a forward + backward at the benchmarked batch (each ncu pass after a plain run of the same command)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 40 > gpurun_out/r02_bench_bf16x3_n1.json 2> gpurun_out/bench_f.err; echo "bench rc=$?"; python tools/show_bench.py gpurun_out/r02_bench_bf16x3_n1.json 2>&1 | tail -24
timeout 900 python bench.py --arch NIPS --steps 40 > gpurun_out/r02_bench_nips_bf16x3_n1.json 2> gpurun_out/bench_fn.err; echo "bench nips rc=$?"; python tools/show_bench.py gpurun_out/r02_bench_nips_bf16x3_n1.json 2>&1 | head -3
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/bench_ref.err; echo "reference rc=$?"; cat gpurun_out/r02_bench_reference_n1.json | cut -c1-300
timeout 300 python bench.py --steps 2 --warmup 3 --no_cpu_baseline --no_e2e --no_variants > gpurun_out/plain_ncu.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no_cpu_baseline --no_e2e --no_variants > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 300 python tools/profile_step.py > gpurun_out/profile_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r02_bf16x3_step python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
