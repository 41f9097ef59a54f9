# final profiles of the round: plain runs first (must exit 0), then the ncu passes of the same commands
timeout 300 python bench.py --steps 2 --warmup 1 --no_cpu_baseline --no_e2e --no_variants > gpurun_out/plain_ncu.log 2>&1; echo plain bench rc=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01_final_launches.csv python bench.py --steps 2 --warmup 1 --no_cpu_baseline --no_e2e --no_variants > gpurun_out/ncu_launch.log 2>&1; echo ncu launches rc=$?
timeout 300 python tools/profile_step.py > gpurun_out/profile_plain.log 2>&1; echo plain profile_step rc=$?
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r01_bf16x3_final6 python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1; echo ncu full rc=$?; tail -2 gpurun_out/ncu_full.log
