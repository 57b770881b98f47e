mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 > gpurun_out/r02_gputest4_full.log; tail -4 gpurun_out/r02_gputest4_full.log; grep -n "expired" gpurun_out/r02_gputest4_full.log | head -3
timeout 300 python profiles/time_kernels.py 8 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_l_attn0_b.csv python profiles/run_stage.py attn 8 0 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_l_attn1_b.csv python profiles/run_stage.py attn 8 1 > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/r02_l_attn0_b.csv | head -6; python profiles/summarize_launches.py gpurun_out/r02_l_attn1_b.csv | head -6
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_c.json 2> gpurun_out/r02_bench_n1_c.err; cat gpurun_out/r02_bench_n1_c.json; tail -3 gpurun_out/r02_bench_n1_c.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 16 > gpurun_out/r02_bench_n1_c16.json 2> gpurun_out/r02_bench_n1_c16.err; cat gpurun_out/r02_bench_n1_c16.json; tail -3 gpurun_out/r02_bench_n1_c16.err
