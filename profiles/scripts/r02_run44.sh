mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_l_train_D_h.csv python bench.py --train D --steps 1 --warmup 0 > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/r02_l_train_D_h.csv 2>/dev/null | head -24
