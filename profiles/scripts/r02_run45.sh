mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q --tb=short 2>&1 | tail -3
timeout 600 python bench.py --train D --steps 5 --warmup 2 --reference-gpu > gpurun_out/r02_train_D_i.json 2> gpurun_out/r02_train_D_i.err; echo "rc=$?"; cut -c1-330 gpurun_out/r02_train_D_i.json; tail -3 gpurun_out/r02_train_D_i.err
timeout 900 python bench.py --train xlarge --steps 2 --warmup 1 > gpurun_out/r02_train_xl_c.json 2> gpurun_out/r02_train_xl_c.err; echo "rc=$?"; cut -c1-330 gpurun_out/r02_train_xl_c.json
