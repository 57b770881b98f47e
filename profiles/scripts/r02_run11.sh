mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 > gpurun_out/r02_gputest7_full.log; tail -3 gpurun_out/r02_gputest7_full.log; grep -n "SI-SDR" gpurun_out/r02_gputest7_full.log | head -12
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_b8_a.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_plain_run.log 2>&1
python profiles/summarize_launches.py gpurun_out/r02_launches_b8_a.csv 2>/dev/null | head -24
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_g.json 2> gpurun_out/r02_bench_n1_g.err; cat gpurun_out/r02_bench_n1_g.json | cut -c1-400; tail -3 gpurun_out/r02_bench_n1_g.err
