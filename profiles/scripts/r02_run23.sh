mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 > gpurun_out/r02_gputest11_full.log; tail -3 gpurun_out/r02_gputest11_full.log
timeout 900 python bench.py --steps 20 --warmup 5 --reference-gpu > gpurun_out/r02_bench_n1_k.json 2> gpurun_out/r02_bench_n1_k.err; cat gpurun_out/r02_bench_n1_k.json | cut -c1-900; tail -3 gpurun_out/r02_bench_n1_k.err
python -c "import __graft_entry__ as g; g.smoke()"
