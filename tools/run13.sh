timeout 300 python tools/profile_step.py > gpurun_out/profile_plain.log 2>&1; echo plain rc=$?
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r01_bf16x3_v5 python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1; echo ncu rc=$?; tail -2 gpurun_out/ncu_full.log
