mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_training.py -m gpu -q --tb=short 2>&1 | grep -v "Warning\|warn\|autocast\|self.gen\|^$\|^tests/\|Consider\|return float" | tail -30
timeout 600 python bench.py --train D --steps 3 --warmup 1 --reference-gpu > gpurun_out/r02_train_D_e.json 2> gpurun_out/r02_train_D_e.err; echo "rc=$?"; cut -c1-400 gpurun_out/r02_train_D_e.json; tail -3 gpurun_out/r02_train_D_e.err
timeout 900 python bench.py --train xlarge --steps 2 --warmup 1 > gpurun_out/r02_train_xl_b.json 2> gpurun_out/r02_train_xl_b.err; echo "rc=$?"; cut -c1-400 gpurun_out/r02_train_xl_b.json; tail -3 gpurun_out/r02_train_xl_b.err
