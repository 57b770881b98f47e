mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2_b.json 2> gpurun_out/r02_bench_n2_b.err; echo "rc=$?"; cut -c1-300 gpurun_out/r02_bench_n2_b.json; tail -2 gpurun_out/r02_bench_n2_b.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-200
