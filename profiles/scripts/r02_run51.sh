mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q --tb=short -s 2>&1 | grep -E "passed|failed|mixed\]|Error|assert" | tail -14
timeout 600 python bench.py --train D --steps 5 --warmup 2 > gpurun_out/r02_train_D_n.json 2> gpurun_out/r02_train_D_n.err; echo "rc=$?"; cut -c1-330 gpurun_out/r02_train_D_n.json; tail -3 gpurun_out/r02_train_D_n.err
