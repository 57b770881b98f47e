mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 > gpurun_out/r02_gputest10_full.log; tail -3 gpurun_out/r02_gputest10_full.log
for ax in 0 1; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_l_attn${ax}_e.csv python profiles/run_stage.py attn 8 $ax > /dev/null 2>&1
echo "axis $ax:"; python profiles/summarize_launches.py gpurun_out/r02_l_attn${ax}_e.csv 2>/dev/null | sed -n 2,5p
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_n1_j.json 2> gpurun_out/r02_bench_n1_j.err; cat gpurun_out/r02_bench_n1_j.json | cut -c1-330; tail -3 gpurun_out/r02_bench_n1_j.err
