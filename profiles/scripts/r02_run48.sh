mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 > gpurun_out/r02_gputest13_full.log; tail -3 gpurun_out/r02_gputest13_full.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_n.json 2> gpurun_out/r02_bench_n1_n.err; cut -c1-300 gpurun_out/r02_bench_n1_n.json; tail -2 gpurun_out/r02_bench_n1_n.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --train D --steps 5 --warmup 2 > gpurun_out/r02_train_D_l.json 2> gpurun_out/r02_train_D_l.err; cut -c1-300 gpurun_out/r02_train_D_l.json
