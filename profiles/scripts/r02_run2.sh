set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_gputest2.log
timeout 300 python profiles/time_kernels.py 8 > gpurun_out/r02_time_kernels_b.txt 2>&1
timeout 300 python profiles/trace_ffn.py > gpurun_out/r02_trace_ffn_b.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err
tail -3 gpurun_out/r02_gputest2.log; cat gpurun_out/r02_time_kernels_b.txt; cat gpurun_out/r02_bench_n1_b.json
