mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 > gpurun_out/r02_gputest6_full.log; tail -3 gpurun_out/r02_gputest6_full.log; grep -n "SI-SDR" gpurun_out/r02_gputest6_full.log | head -30
for ax in 0 1; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_l_attn${ax}_d.csv python profiles/run_stage.py attn 8 $ax > /dev/null 2>&1
echo "axis $ax:"; python profiles/summarize_launches.py gpurun_out/r02_l_attn${ax}_d.csv 2>/dev/null | sed -n 2,5p
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_f.json 2> gpurun_out/r02_bench_n1_f.err; cat gpurun_out/r02_bench_n1_f.json | cut -c1-400; tail -3 gpurun_out/r02_bench_n1_f.err
