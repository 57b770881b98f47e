mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 > gpurun_out/r02_gputest12_full.log; tail -3 gpurun_out/r02_gputest12_full.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_m.json 2> gpurun_out/r02_bench_n1_m.err; cut -c1-600 gpurun_out/r02_bench_n1_m.json; tail -2 gpurun_out/r02_bench_n1_m.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
