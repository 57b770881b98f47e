// tcgen05 / TMEM kernels (TFL_PRECISION_BF16).  Placeholder until the first GPU bring-up
// of the fp32 path is green: every entry point fails loudly, there is no fallback.
#pragma once
#include "common.cuh"

namespace tfl {
inline size_t tc_ffn_image_bytes(int C, int H, int K) { return 0; }
inline int tc_pack_ffn(const float*, const float*, const float*, const float*, char*, int, int, int, cudaStream_t) { return 0; }
inline size_t tc_workspace_bytes(const tfl_plan*, int, int, int) { return 0; }
inline int tc_ffn(const tfl_plan*, const char*, int, int, int, float*, int, int, int, char*, cudaStream_t) {
  set_error("bf16 tcgen05 FFN kernel not built");
  return -3;
}
inline int tc_attn(const tfl_plan*, const char*, int, int, float*, int, int, int, char*, char*, size_t, size_t, size_t, cudaStream_t) {
  set_error("bf16 tcgen05 attention kernel not built");
  return -3;
}
inline int tc_path_forward(const tfl_plan*, const char*, int, int, float*, int, int, int, char*, char*, size_t, size_t, size_t, cudaStream_t) {
  set_error("bf16 tcgen05 path not built");
  return -3;
}
}  // namespace tfl
