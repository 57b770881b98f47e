mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 > gpurun_out/r02_gputest8_full.log; tail -3 gpurun_out/r02_gputest8_full.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_h.json 2> gpurun_out/r02_bench_n1_h.err; cat gpurun_out/r02_bench_n1_h.json | cut -c1-330; tail -3 gpurun_out/r02_bench_n1_h.err
TFL_NO_PDL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_h_nopdl.json 2> gpurun_out/r02_bench_n1_h_nopdl.err; cat gpurun_out/r02_bench_n1_h_nopdl.json | cut -c1-330; tail -3 gpurun_out/r02_bench_n1_h_nopdl.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_h2.json 2> gpurun_out/r02_bench_n1_h2.err; cat gpurun_out/r02_bench_n1_h2.json | cut -c1-330
