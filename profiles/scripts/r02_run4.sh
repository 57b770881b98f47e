mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,temperature.gpu,power.draw --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 > gpurun_out/r02_gputest3_full.log; tail -5 gpurun_out/r02_gputest3_full.log; grep -n "expired" gpurun_out/r02_gputest3_full.log | head -3
timeout 300 python profiles/time_kernels.py 8 2>&1 | tail -4 | tee gpurun_out/r02_time_kernels_b.txt
timeout 300 python profiles/trace_ffn.py > gpurun_out/r02_trace_ffn_b.txt 2>&1; tail -3 gpurun_out/r02_trace_ffn_b.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err; cat gpurun_out/r02_bench_n1_b.json; tail -3 gpurun_out/r02_bench_n1_b.err
