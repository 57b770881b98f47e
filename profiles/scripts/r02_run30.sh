mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q --tb=short -k "two_gpus or cfg1" 2>&1 | grep -v "Warning\|warn\|autocast\|self.gen\|^$\|^tests/" | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --train D --steps 3 --warmup 1 > gpurun_out/r02_train_D_n2.json 2> gpurun_out/r02_train_D_n2.err; echo "rc=$?"; cut -c1-1200 gpurun_out/r02_train_D_n2.json; tail -3 gpurun_out/r02_train_D_n2.err
timeout 900 python bench.py --train xlarge --steps 2 --warmup 1 > gpurun_out/r02_train_xl.json 2> gpurun_out/r02_train_xl.err; echo "rc=$?"; cut -c1-1500 gpurun_out/r02_train_xl.json; tail -3 gpurun_out/r02_train_xl.err
