mkdir -p gpurun_out
for ax in 0 1; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_l_attn${ax}_f.csv python profiles/run_stage.py attn 8 $ax > /dev/null 2>&1
echo "axis $ax: $(python profiles/summarize_launches.py gpurun_out/r02_l_attn${ax}_f.csv 2>/dev/null | grep attn_tc2)"
done
timeout 900 python -m pytest tests -m gpu -x -q -k "not fullsize" 2>&1 | tail -3
