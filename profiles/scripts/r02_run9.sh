mkdir -p gpurun_out
V=mss_tf_locoformer_b200/csrc/variants
timeout 900 python -m pytest tests -m gpu -x -q -k "not fullsize" 2>&1 | tail -3
for ax in 0 1; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_l_attn${ax}_c.csv python profiles/run_stage.py attn 8 $ax > /dev/null 2>&1
TFL_LIB=$V/lib_nopf.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_l_attn${ax}_c_nopf.csv python profiles/run_stage.py attn 8 $ax > /dev/null 2>&1
echo "axis $ax prefetch:"; python profiles/summarize_launches.py gpurun_out/r02_l_attn${ax}_c.csv 2>/dev/null | sed -n 2,5p
echo "axis $ax no prefetch:"; python profiles/summarize_launches.py gpurun_out/r02_l_attn${ax}_c_nopf.csv 2>/dev/null | sed -n 2,5p
done
timeout 300 python profiles/time_kernels.py 8 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_e.json 2> gpurun_out/r02_bench_n1_e.err; cat gpurun_out/r02_bench_n1_e.json | cut -c1-400; tail -3 gpurun_out/r02_bench_n1_e.err
