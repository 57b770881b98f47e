mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q --tb=short -k "widths or attention_backward" 2>&1 | tail -2
timeout 600 python bench.py --train D --steps 5 --warmup 2 > gpurun_out/r02_train_D_j.json 2> gpurun_out/r02_train_D_j.err; echo "rc=$?"; cut -c1-330 gpurun_out/r02_train_D_j.json; tail -3 gpurun_out/r02_train_D_j.err
