mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_training.py -m gpu -q --tb=short -s 2>&1 | grep -v "Warning\|warn\|autocast\|self.gen\|^$" | tail -40
timeout 600 python bench.py --train D --steps 3 --warmup 1 --reference-gpu > gpurun_out/r02_train_D.json 2> gpurun_out/r02_train_D.err; echo "rc=$?"; cut -c1-1500 gpurun_out/r02_train_D.json; tail -3 gpurun_out/r02_train_D.err
