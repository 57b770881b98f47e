mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 > gpurun_out/r02_gputest9_full.log; tail -3 gpurun_out/r02_gputest9_full.log
timeout 600 python bench.py --steps 20 --warmup 5 --reference-gpu > gpurun_out/r02_bench_n1_i.json 2> gpurun_out/r02_bench_n1_i.err; cat gpurun_out/r02_bench_n1_i.json | cut -c1-330; tail -3 gpurun_out/r02_bench_n1_i.err
timeout 600 python bench.py --steps 10 --warmup 3 --variant Y --no-cpu-baseline > gpurun_out/r02_bench_y_b.json 2> gpurun_out/r02_bench_y_b.err; cat gpurun_out/r02_bench_y_b.json | cut -c1-330
timeout 600 python bench.py --steps 10 --warmup 3 --model bs > gpurun_out/r02_bench_bs_b.json 2> gpurun_out/r02_bench_bs_b.err; cat gpurun_out/r02_bench_bs_b.json | cut -c1-330
timeout 600 python bench.py --steps 5 --warmup 3 --precision fp32 --no-cpu-baseline > gpurun_out/r02_bench_fp32.json 2> gpurun_out/r02_bench_fp32.err; cat gpurun_out/r02_bench_fp32.json | cut -c1-330
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ffn_tc2|attn_tc2|qkv_tc|proj_tc|attn_tail' -c 5 -o gpurun_out/r02_hot python profiles/run_both.py > gpurun_out/r02_ncu_hot.log 2>&1; tail -2 gpurun_out/r02_ncu_hot.log
timeout 600 ncu --set full --clock-control none -k regex:'stft_kernel|enc_conv|gln_|dec_conv|istft_ola' -o gpurun_out/r02_hbm_b python profiles/run_forward.py > gpurun_out/r02_hbm_ncu_b.log 2>&1; tail -2 gpurun_out/r02_hbm_ncu_b.log
