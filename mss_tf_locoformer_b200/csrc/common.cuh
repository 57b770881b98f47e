// Shared declarations for the TF-Locoformer sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/tfl.h"

namespace tfl {

void set_error(const char* fmt, ...);

#define TFL_CHECK(cond, ...)                         \
  do {                                               \
    if (!(cond)) {                                   \
      ::tfl::set_error(__VA_ARGS__);                 \
      return -1;                                     \
    }                                                \
  } while (0)

#define TFL_CUDA(expr)                                                                  \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      ::tfl::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
      return -2;                                                                        \
    }                                                                                   \
  } while (0)

extern std::atomic<unsigned long long> g_launches;  // kernels launched by this library since load (bench.py's gpu_launches)
#define TFL_LAUNCH_CHECK()                                        \
  do {                                                            \
    ::tfl::g_launches.fetch_add(1, std::memory_order_relaxed);    \
    TFL_CUDA(cudaGetLastError());                                 \
  } while (0)

// Opt in to > 48 KB of dynamic shared memory.  The attribute belongs to the (device, function) pair, so it is set on
// every launch of the calling thread's current device instead of being cached per thread (a host-only driver call).
template <class Kernel>
inline cudaError_t opt_in_smem(Kernel kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// Programmatic dependent launch for the back-to-back persistent kernels of a step: the next kernel's CTAs become
// resident as the previous kernel's CTAs exit and run their prologue (barrier init, TMEM allocation, table loads) while
// the tail of the previous kernel is still draining; they block in griddep_wait() before touching anything the
// previous kernel wrote.  tfl_debug_set_option(TFL_OPT_PDL, 0) launches plainly (A/B).
int tfl_option(int key);
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tfl_option(4 /* TFL_OPT_PDL */) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// Maps (sequence s, position p) of one Locoformer path onto the channels-last residual
// stream x[B, Tf, F, C] without materialising the reference's transposes
// (models/mss_tflocoformer.py:339-344): element offset = (s / inner) * outer_stride
// + (s % inner) * inner_stride + p * pos_stride.
struct SeqMap {
  int inner;
  long long outer_stride, inner_stride, pos_stride;
  __host__ __device__ __forceinline__ long long base(int s) const {
    return (long long)(s / inner) * outer_stride + (long long)(s % inner) * inner_stride;
  }
};

inline SeqMap make_seq_map(int axis, int n_frames, int n_freq, int C) {
  SeqMap m;
  if (axis == TFL_AXIS_FREQ) {  // s = b*Tf + t, p = f
    m.inner = 1 << 30; m.outer_stride = 0; m.inner_stride = (long long)n_freq * C; m.pos_stride = C;
  } else {                      // s = b*F + f, p = t
    m.inner = n_freq; m.outer_stride = (long long)n_frames * n_freq * C; m.inner_stride = C;
    m.pos_stride = (long long)n_freq * C;
  }
  return m;
}

// 128-row tiling of the (sequence, position) rows of one attention call for the q|k|v and head-merge GEMMs.
//   legacy  : NTL = ceil(L / 128) tiles per sequence, the last one partial;
//   packed  : when only a few rows spill over a tile boundary (1025 = 8 * 128 + 1, 259 = 2 * 128 + 3) the NTF = L / 128 full
//             tiles of every sequence come first (tile = s * NTF + jt) and the spill-over rows of ALL sequences are packed
//             into ceil(nseq * n_tail / 128) tiles behind them -- instead of one nearly empty tile per sequence (a third
//             of the time-axis tiles of a 6-s segment).
struct TileMap {
  int nseq, L, NTL, NTF, n_tail, packed;
  int n_full;    // nseq * NTF (packed) or nseq * NTL (legacy)
  int n_tiles;
  __host__ __device__ __forceinline__ bool locate(int tile, int m, int& s, int& j) const {
    if (!packed) { s = tile / NTL; j = (tile - s * NTL) * 128 + m; return j < L; }
    if (tile < n_full) { s = tile / NTF; j = (tile - s * NTF) * 128 + m; return true; }
    const long long g = (long long)(tile - n_full) * 128 + m;
    s = (int)(g / n_tail); j = NTF * 128 + (int)(g - (long long)s * n_tail);
    return s < nseq;
  }
  // tile of the attention-output image that holds row j of sequence s, and the row inside it
  __host__ __device__ __forceinline__ void o_slot(int s, int j, long long& tile, int& row) const {
    if (!packed) { tile = (long long)s * NTL + (j >> 7); row = j & 127; return; }
    if (j < NTF * 128) { tile = (long long)s * NTF + (j >> 7); row = j & 127; return; }
    const long long g = (long long)s * n_tail + (j - NTF * 128);
    tile = n_full + (g >> 7); row = (int)(g & 127);
  }
};
inline TileMap make_tile_map(int nseq, int L, bool pack_tail) {
  TileMap t;
  t.nseq = nseq; t.L = L; t.NTL = (L + 127) / 128; t.NTF = L / 128; t.n_tail = L - t.NTF * 128;
  t.packed = (pack_tail && t.n_tail > 0 && t.NTF > 0) ? 1 : 0;
  if (t.packed) {
    t.n_full = nseq * t.NTF;
    t.n_tiles = t.n_full + (int)(((long long)nseq * t.n_tail + 127) / 128);
  } else {
    t.n_full = nseq * t.NTL;
    t.n_tiles = t.n_full;
  }
  return t;
}

inline SeqMap make_dense_map(long long seq_stride, long long pos_stride) {
  SeqMap m; m.inner = 1 << 30; m.outer_stride = 0; m.inner_stride = seq_stride; m.pos_stride = pos_stride;
  return m;
}

constexpr int ROPE_TAB_LEN = 2304;   // positions covered by the pack-time RoPE table (F = 2049 of n_fft 4096 fits); longer: per call
// ---- packed weight image -------------------------------------------------------------
struct FfnPack {
  size_t gamma;      // [C] fp32
  size_t w1;         // fp32 [K][C][2H] with (value, gate) column-interleaved
  size_t b1;         // [2H] interleaved
  size_t b1raw;      // [2H] reference order (value | gate), used by the tcgen05 epilogue
  size_t w2;         // fp32 [K][H][C], tap order reversed (tap k' multiplies g[i + k'])
  size_t b2;         // [C]
  size_t tc;         // bf16 tcgen05 image (0 = none)
  size_t tc2;        // bf16 image for the 2-CTA kernel (halves per CTA); valid when tc2_ok
  int tc2_ok;
  int hidden;
};
struct PathPack {
  FfnPack ffn[2];
  size_t attn_gamma, wqkv /*[C][3A]*/, wo /*[A][C]*/, rope /*[hd/2]*/;
  size_t tc_qkv, tc_wo;
  size_t rope_tab;   // (cos, sin) of positions 0 .. ROPE_TAB_LEN - 1, frequency-major [HDP/2][ROPE_TAB_LEN] float2 (bf16 path)
};
struct PackLayout {
  size_t enc_w /*[3][3][Cin][C]*/, enc_b, gln_w, gln_b, dec_w /*[9 taps][8 outputs][C]*/, dec_b;
  size_t twiddle /*[n_fft/2] float2*/, window /*[n_fft]*/;
  std::vector<PathPack> paths;  // [2*layer + axis]
  size_t total;
};

}  // namespace tfl

struct tfl_plan {
  tfl_config cfg;
  tfl::PackLayout lay;
  int head_dim;
  int n_ffn;
  int sm_count;
};
