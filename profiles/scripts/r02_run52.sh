mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 --train xlarge --steps 2 --warmup 1 > gpurun_out/r02_train_xl_n8.json 2> gpurun_out/r02_train_xl_n8.err; echo "rc=$?"; cut -c1-330 gpurun_out/r02_train_xl_n8.json; tail -2 gpurun_out/r02_train_xl_n8.err
