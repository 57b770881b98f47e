mkdir -p gpurun_out
V=mss_tf_locoformer_b200/csrc/variants
for lib in diag1 diag2; do
for ax in 0 1; do
export TFL_LIB=$V/lib_$lib.so
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_diag_${lib}_$ax.csv python profiles/run_stage.py attn 8 $ax > /dev/null 2>&1
echo "$lib axis $ax: $(python profiles/summarize_launches.py gpurun_out/r02_diag_${lib}_$ax.csv 2>/dev/null | grep attn_tc2)"
done; done
