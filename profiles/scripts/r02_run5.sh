mkdir -p gpurun_out
V=mss_tf_locoformer_b200/csrc/variants
echo "== default (no prefetch, setmaxnreg, merged FULL, qkv UF 8)"; timeout 200 python profiles/time_kernels.py 8 2>&1 | tail -3
for n in pf1 unmerged pf0_nosm uf4; do echo "== $n"; TFL_LIB=$V/lib_$n.so timeout 200 python profiles/time_kernels.py 8 2>&1 | tail -3; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_l_attn0.csv python profiles/run_stage.py attn 8 0 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_l_attn1.csv python profiles/run_stage.py attn 8 1 > /dev/null 2>&1
TFL_LIB=$V/lib_pf1.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_l_attn0_pf1.csv python profiles/run_stage.py attn 8 0 > /dev/null 2>&1
python profiles/summarize_launches.py gpurun_out/r02_l_attn0.csv; python profiles/summarize_launches.py gpurun_out/r02_l_attn1.csv; python profiles/summarize_launches.py gpurun_out/r02_l_attn0_pf1.csv
timeout 300 python profiles/trace_ffn.py > gpurun_out/r02_trace_ffn_c.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ffn_tc2|attn_tc2|qkv_tc' -c 3 -o gpurun_out/r02_ffn_attn python profiles/run_both.py > gpurun_out/r02_ncu_both.log 2>&1
tail -2 gpurun_out/r02_ncu_both.log
