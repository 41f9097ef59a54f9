timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench rc=$?; tail -2 gpurun_out/bench_n1.err
python tools/show_bench.py gpurun_out/bench_n1.json 2>&1 | tail -24
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_n1.json 2> gpurun_out/bench_ref.err; echo ref rc=$?; cut -c1-400 gpurun_out/bench_reference_n1.json
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
