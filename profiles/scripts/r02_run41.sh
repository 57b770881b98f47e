mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q --tb=short -s -k "xlarge_widths" 2>&1 | grep -E "passed|failed|widths|Error|assert" | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:tap_wgrad_mma|tap_gemm_mma|attn_bwd_dkv_mma" -c 12 -o gpurun_out/r02_train_kernels python bench.py --train D --steps 1 --warmup 0 > gpurun_out/ncu_train.log 2>&1; tail -2 gpurun_out/ncu_train.log
