mkdir -p gpurun_out
# A/B/C of the attention kernel: exponentials of 0 / 2 / 4 of every 8 pairs on the FMA pipe (ATT2_POLY_MASK), same box;
# then the bf16 tcgen05 tests and one full-size parity test on the fastest build
V=mss_tf_locoformer_b200/csrc
best=""; bestsum=1000000
for lib in libtfl_b200.so libtfl_b200_p0x88.so libtfl_b200_p0xAA.so; do
  TFL_LIB=$PWD/$V/$lib timeout 40 python profiles/time_attn.py 8 > gpurun_out/r02_attn_poly_$lib.txt 2>&1
  grep "kernel v2" gpurun_out/r02_attn_poly_$lib.txt | sed "s/^/$lib: /"
  sum=$(grep "kernel v2" gpurun_out/r02_attn_poly_$lib.txt | awk '{s+=$5} END {printf "%d", s*1000}')
  if [ -n "$sum" ] && [ "$sum" -gt 0 ] && [ "$sum" -lt "$bestsum" ]; then bestsum=$sum; best=$lib; fi
done
echo "fastest: $best ($bestsum us for the two axes)"
echo "$best" > gpurun_out/r02_attn_poly_choice.txt
TFL_LIB=$PWD/$V/$best timeout 85 python -m pytest tests/test_gpu_tcgen05.py tests/test_gpu_fullsize.py::test_variant_d_full_depth_6s_bf16 -x -q -m gpu 2>&1 | tail -3
