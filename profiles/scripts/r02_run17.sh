V=mss_tf_locoformer_b200/csrc/variants
for i in 1 2 3; do
timeout 200 python profiles/ab_time.py 8 2>&1 | tail -1
TFL_LIB=$V/lib_prev.so timeout 200 python profiles/ab_time.py 8 2>&1 | tail -1
done
