mkdir -p gpurun_out
V=mss_tf_locoformer_b200/csrc/variants
for lib in default noearly default noearly; do
for ax in 0; do
if [ $lib = default ]; then unset TFL_LIB; else export TFL_LIB=$V/lib_$lib.so; fi
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ffn_${lib}_$ax.csv python profiles/run_stage.py ffn 8 $ax > /dev/null 2>&1
echo "$lib axis $ax: $(python profiles/summarize_launches.py gpurun_out/r02_ffn_${lib}_$ax.csv 2>/dev/null | grep ffn_tc2)"
done; done
unset TFL_LIB
timeout 900 python -m pytest tests -m gpu -x -q -k "tcgen05 or golden or real_width" 2>&1 | tail -3
timeout 300 python profiles/trace_ffn.py 8 300 > gpurun_out/r02_trace_ffn_b8_q300_c.txt 2>&1; sed -n 4,14p gpurun_out/r02_trace_ffn_b8_q300_c.txt
