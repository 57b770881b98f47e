mkdir -p gpurun_out
# ncu --set full of the two kernels rewritten at the end of round 2 (scatter-form decoder, half-size radix-4 iSTFT)
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'dec_conv_scatter|istft_ola' -c 2 -o gpurun_out/r02_hbm_c python profiles/run_forward.py > gpurun_out/r02_ncu_hbm_c.log 2>&1
tail -2 gpurun_out/r02_ncu_hbm_c.log
# launch list of the bench command with the final build
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_b8_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_bench_final.log 2>&1
python profiles/summarize_launches.py gpurun_out/r02_launches_bench_b8_final.csv | head -24
