mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tcgen05.py -m gpu -x -q -k attention 2>&1 | tail -2
for cfg in "0 6:1" "1 6:2" "0 6:0"; do
set -- $cfg
TFL_OPTS=$2 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_l_attn$1_fold_$2.csv python profiles/run_stage.py attn 8 $1 > /dev/null 2>&1
echo "axis $1 opts $2:"; python profiles/summarize_launches.py gpurun_out/r02_l_attn$1_fold_$2.csv 2>/dev/null | sed -n 2,5p
done
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_n1_fold2.json 2> gpurun_out/r02_bench_n1_fold2.err; cut -c1-300 gpurun_out/r02_bench_n1_fold2.json; tail -2 gpurun_out/r02_bench_n1_fold2.err
