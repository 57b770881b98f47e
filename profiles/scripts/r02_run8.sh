mkdir -p gpurun_out
V=mss_tf_locoformer_b200/csrc/variants
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 > gpurun_out/r02_gputest5_full.log; tail -3 gpurun_out/r02_gputest5_full.log; grep -n "expired" gpurun_out/r02_gputest5_full.log | head -3
echo "== default (CTA-scope waits + releases)"; timeout 200 python profiles/time_kernels.py 8 2>&1 | tail -3
for n in clrel clwait; do echo "== $n"; TFL_LIB=$V/lib_$n.so timeout 200 python profiles/time_kernels.py 8 2>&1 | tail -3; done
echo "== default again"; timeout 200 python profiles/time_kernels.py 8 2>&1 | tail -3
timeout 300 python profiles/trace_ffn.py 8 300 > gpurun_out/r02_trace_ffn_b8_q300_b.txt 2>&1; sed -n 4,16p gpurun_out/r02_trace_ffn_b8_q300_b.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_d.json 2> gpurun_out/r02_bench_n1_d.err; cat gpurun_out/r02_bench_n1_d.json | cut -c1-1200; tail -3 gpurun_out/r02_bench_n1_d.err

