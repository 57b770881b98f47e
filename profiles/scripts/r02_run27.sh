mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tcgen05.py -m gpu -x -q 2>&1 | tail -5
for ax in 0 1; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_l_attn${ax}_pack.csv python profiles/run_stage.py attn 8 $ax > /dev/null 2>&1
echo "axis $ax:"; python profiles/summarize_launches.py gpurun_out/r02_l_attn${ax}_pack.csv 2>/dev/null | sed -n 1,8p
done
