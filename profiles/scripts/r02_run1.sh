set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/r02_gputest1.log
timeout 600 python bench.py --steps 10 --warmup 3 --reference-gpu > gpurun_out/r02_bench_n1_a.json 2> gpurun_out/r02_bench_n1_a.err
timeout 300 python profiles/trace_ffn.py > gpurun_out/r02_trace_ffn_a.txt 2>&1
timeout 300 python profiles/time_kernels.py 8 > gpurun_out/r02_time_kernels_a.txt 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --variant Y --no-cpu-baseline > gpurun_out/r02_bench_y_a.json 2> gpurun_out/r02_bench_y_a.err
timeout 300 python bench.py --steps 5 --warmup 3 --model bs > gpurun_out/r02_bench_bs_a.json 2> gpurun_out/r02_bench_bs_a.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'stft_kernel|enc_conv|gln_apply|rms_group_norm|dec_conv|istft_ola' -o gpurun_out/r02_hbm python profiles/run_forward.py > gpurun_out/r02_hbm_ncu.log 2>&1
ls -la gpurun_out
