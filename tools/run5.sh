for s in 1 2 4; do timeout 600 python bench.py --no_cpu_baseline --e2e_slices $s > gpurun_out/bench_e2e_s$s.log 2> gpurun_out/bench_e2e_s$s.err; echo bench s=$s rc=$?; done
