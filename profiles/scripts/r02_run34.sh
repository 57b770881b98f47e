mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tcgen05.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -4
for ax in 0 1; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_l_attn${ax}_fold.csv python profiles/run_stage.py attn 8 $ax > /dev/null 2>&1
echo "axis $ax:"; python profiles/summarize_launches.py gpurun_out/r02_l_attn${ax}_fold.csv 2>/dev/null | sed -n 1,6p
done
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_n1_fold.json 2> gpurun_out/r02_bench_n1_fold.err; cut -c1-300 gpurun_out/r02_bench_n1_fold.json; tail -2 gpurun_out/r02_bench_n1_fold.err
