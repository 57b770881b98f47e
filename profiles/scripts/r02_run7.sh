mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpus or cli or espnet or device_metrics" 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "rc=$?"; cat gpurun_out/r02_bench_n2.json; tail -3 gpurun_out/r02_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --track 240 > gpurun_out/r02_track_n2.json 2> gpurun_out/r02_track_n2.err; echo "rc=$?"; cat gpurun_out/r02_track_n2.json; tail -3 gpurun_out/r02_track_n2.err
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --track 240 > gpurun_out/r02_track_n1.json 2> gpurun_out/r02_track_n1.err; cat gpurun_out/r02_track_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>&1 | tail -2 | cut -c1-400
timeout 300 python profiles/trace_ffn.py 8 300 > gpurun_out/r02_trace_ffn_b8_q300.txt 2>&1; tail -3 gpurun_out/r02_trace_ffn_b8_q300.txt
