// tcgen05 / TMEM kernels of TFL_PRECISION_BF16.
//
// K4  ffn_tc_kernel: y = x + ConvSwiGLU(RMSGroupNorm(x)) along one axis as ONE persistent,
//     warp-specialised kernel (models/mss_tflocoformer.py:443-447,459-462 -> :626-655):
//       producers   fp32 residual rows -> RMSGroupNorm -> bf16 chunk-major A tile in smem
//       loader      weight stages (pre-packed bf16 smem images) via 1-D bulk async copies
//       MMA thread  conv1d as KT row-shifted tcgen05.mma taps into TMEM (value | gate halves),
//                   transposed conv as KT taps over the SwiGLU'd hidden tile
//       epilogue    TMEM -> bias + SwiGLU -> bf16 hidden tile in smem (never touches HBM);
//                   final: TMEM -> + bias + residual -> y (fp32)
//     Sequences are laid on a "stream" with period P = S + KT - 1 rows (KT - 1 shared zero rows
//     between consecutive sequences == the reference's zero padding AFTER the norm, :640-644),
//     so M tiles run across sequence boundaries and both axes use the same kernel through SeqMap.
//     Each CTA works on NT = 2 M-tiles at once so every weight stage fetched from L2 feeds 256 rows.
//
//     TMEM (512 columns): D1[t] = 128 columns (64 value | 64 gate) per tile, D2[t] = C columns.
//     Tensor-pipe order per hidden chunk c: M1[0](c) M1[1](c) M2[0](c-1) M2[1](c-1); the SwiGLU
//     epilogue of chunk c overlaps M2(c-1) and M1(c+1).
#pragma once
#include <nvtx3/nvToolsExt.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace tfl {

constexpr int TC_HC = 64;             // hidden channels per chunk (D1 tile = 2*HC TMEM columns)
constexpr int TC_SMEM_MAX = 232448;   // 227 KB opt-in shared memory per CTA
#ifndef TC_A_EXTRA
#define TC_A_EXTRA 2                  // A-tile slots beyond the NT tiles being multiplied (prefetch depth)
#endif

struct FfnTcGeom {
  int C, H, KT, G, NT, NS;            // NS = weight ring stages
  int AR;                             // rows per A / G tile = 128 + KT - 1
  int TS;                             // output rows per tile = 128 - (KT - 1)
  int NC;                             // hidden chunks
  int KH;                             // K-halves per W1 tap stage (1 or 2)
  int TPS, KS;                        // taps per W2 stage, W2 stages per chunk
  uint32_t stage_bytes, a_slot_bytes, g_buf_bytes;
  uint32_t off_a, off_g, off_w, off_tab, off_bar, smem_bytes;
  int threads;
};

inline bool ffn_tc_geometry(int C, int H, int KT, int G, FfnTcGeom* g) {
  if (C % 16 != 0 || C > 256 || H % TC_HC != 0 || (C / G) % 4 != 0 || KT < 1 || KT > 8) return false;
  g->C = C; g->H = H; g->KT = KT; g->G = G;
  g->NT = C <= 128 ? 2 : 1;
  g->AR = 128 + KT - 1; g->TS = 128 - (KT - 1);
  g->NC = H / TC_HC;
  g->KH = (C % 32 == 0) ? 2 : 1;
  g->TPS = 2 / g->KH; g->KS = (KT + g->TPS - 1) / g->TPS;
  g->stage_bytes = 256u * C / g->KH;
  g->a_slot_bytes = (uint32_t)(C / 8) * g->AR * 16;
  g->g_buf_bytes = (uint32_t)(TC_HC / 8) * g->AR * 16;
  uint32_t off = 0;
  g->off_a = off; off += (g->NT == 2 ? g->NT + TC_A_EXTRA : 2) * g->a_slot_bytes;
  g->off_g = off; off += g->NT * g->g_buf_bytes;
  g->off_tab = off; off += (2 * H + 2 * C) * 4;
  off = (off + 15) & ~15u;
  g->off_bar = off; off += 512;
  off = (off + 127) & ~127u;
  g->off_w = off;
  if (off + 2 * g->stage_bytes > (uint32_t)TC_SMEM_MAX) return false;
  g->NS = (int)((TC_SMEM_MAX - off) / g->stage_bytes);
  if (g->NS > 8) g->NS = 8;
  g->smem_bytes = off + g->NS * g->stage_bytes;
  g->threads = 32 * (6 + 5 * g->NT);   // loader, NT + 1 MMA warps, 4 producers, 4 * NT epilogue warps
  return true;
}

inline size_t tc_ffn_image_bytes(int C, int H, int K) {
  FfnTcGeom g;
  if (!ffn_tc_geometry(C, H, K, 1, &g)) return 0;
  return (size_t)g.NC * (K * g.KH + g.KS) * g.stage_bytes;
}

// Weight image: first, per hidden chunk c, the KT*KH conv1d stages W1(c, k, half) (B operand [128 rows = 64 value |
// 64 gate] x [C/KH input channels], chunk-major); then, per chunk, the KS transposed-conv stages W2(c, s)
// (TPS taps x [C rows] x [64 hidden channels]).  The loader streams W1(q) followed by W2(q - 1) for the running
// chunk index q, across tile-pair boundaries.
__global__ void tc_pack_ffn_kernel(const float* __restrict__ w1, const float* __restrict__ w2,
                                   __nv_bfloat16* __restrict__ img, int C, int H, int KT, int KH, int TPS, int KS) {
  const int NC = H / TC_HC, n1 = KT * KH, n_per = n1 + KS;
  const int CK = C / KH;                                   // input channels per W1 stage
  const long long stage_elems = 128LL * CK;
  const long long total = (long long)NC * n_per * stage_elems;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx / stage_elems);
    const int e = (int)(idx % stage_elems);
    int is_w1, c, sub;
    if (g < NC * n1) { is_w1 = 1; c = g / n1; sub = g - c * n1; }
    else { is_w1 = 0; c = (g - NC * n1) / KS; sub = (g - NC * n1) - c * KS; }
    (void)n_per;
    float v = 0.f;
    if (is_w1) {
      const int k = sub / KH, hf = sub % KH;
      const int chunk = e / (128 * 8), n = (e / 8) % 128, cc = hf * CK + chunk * 8 + (e & 7);
      const int row = n < TC_HC ? c * TC_HC + n : H + c * TC_HC + (n - TC_HC);
      v = w1[((size_t)row * C + cc) * KT + k];
    } else {
      const int per_tap = (TC_HC / 8) * C * 8;
      const int tl = e / per_tap, e2 = e % per_tap;
      const int chunk = e2 / (C * 8), n = (e2 / 8) % C, hh = chunk * 8 + (e2 & 7);
      const int tap = TPS * sub + tl;
      if (tl < TPS && tap < KT) v = w2[((size_t)(c * TC_HC + hh) * C + n) * KT + (KT - 1 - tap)];
    }
    img[idx] = __float2bfloat16_rn(v);
  }
}

inline int tc_pack_ffn(const float* w1, const float* b1, const float* w2, const float* b2, char* img, int C, int H,
                       int K, cudaStream_t st) {
  (void)b1; (void)b2;
  FfnTcGeom g;
  if (!ffn_tc_geometry(C, H, K, 1, &g)) return 0;  // shape not covered by the tcgen05 path
  tc_pack_ffn_kernel<<<592, 256, 0, st>>>(w1, w2, (__nv_bfloat16*)img, C, H, K, g.KH, g.TPS, g.KS);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

struct FfnTcParams {
  const float* x;         // residual stream in (read by the norm prologue and the residual add)
  float* y;               // residual stream out; MUST NOT alias x: neighbouring tiles read x halo rows
  SeqMap map;
  int S, P;               // sequence length, stream period S + KT - 1
  long long R;            // stream rows = nseq * P
  int n_tiles;
  const float* gamma; const float* b1 /*raw [2H]*/; const float* b2;
  const char* img;        // weight image
  float eps;
};

// value * SiLU(gate) with one MUFU op: sigmoid(g) = 0.5 + 0.5 * tanh(g / 2)  (tanh.approx: rel. error ~2^-11,
// far below the bf16 rounding of the result)
__device__ __forceinline__ float swiglu_fast(float value, float gate) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * gate));
  return value * gate * fmaf(0.5f, t, 0.5f);
}

// one tcgen05.mma from 32-bit descriptor halves (keeps the issue loop to a handful of integer adds)
__device__ __forceinline__ void mma_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 ad, bd;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 ad, {%1, %2};\n\t"
      "mov.b64 bd, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// N back-to-back tcgen05.mma whose A / B descriptors advance by a fixed step: one asm block, so the descriptor
// arithmetic stays in the uniform datapath instead of a register -> uniform-register transfer per MMA.
template <int N>
__device__ __forceinline__ void mma_burst(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                          uint32_t acc_first, uint32_t a_step, uint32_t b_step) {
  static_assert(N == 1 || N == 2 || N == 4, "burst length");
  if constexpr (N == 4) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 a1, a2, a3, b1, b2, b3;\n\t.reg .b64 ad, bd;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "add.u32 a1, %1, %6;\n\tadd.u32 a2, a1, %6;\n\tadd.u32 a3, a2, %6;\n\t"
        "add.u32 b1, %2, %7;\n\tadd.u32 b2, b1, %7;\n\tadd.u32 b3, b2, %7;\n\t"
        "mov.b64 ad, {%1, %3};\n\tmov.b64 bd, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %4, p;\n\t"
        "mov.b64 ad, {a1, %3};\n\tmov.b64 bd, {b1, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %4, 1;\n\t"
        "mov.b64 ad, {a2, %3};\n\tmov.b64 bd, {b2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %4, 1;\n\t"
        "mov.b64 ad, {a3, %3};\n\tmov.b64 bd, {b3, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %4, 1;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc_first), "r"(a_step), "r"(b_step)
        : "memory");
  } else if constexpr (N == 2) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 a1, b1;\n\t.reg .b64 ad, bd;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "add.u32 a1, %1, %6;\n\tadd.u32 b1, %2, %7;\n\t"
        "mov.b64 ad, {%1, %3};\n\tmov.b64 bd, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %4, p;\n\t"
        "mov.b64 ad, {a1, %3};\n\tmov.b64 bd, {b1, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], ad, bd, %4, 1;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc_first), "r"(a_step), "r"(b_step)
        : "memory");
  } else {
    mma_lohi(d_tmem, a_lo, hi, b_lo, hi, idesc, acc_first);
  }
}
// n MMAs (any n) as bursts of 4 / 2 / 1
__device__ __forceinline__ void mma_run(int n, uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                        uint32_t acc_first, uint32_t a_step, uint32_t b_step) {
  while (n >= 4) { mma_burst<4>(d_tmem, a_lo, b_lo, hi, idesc, acc_first, a_step, b_step); a_lo += 4 * a_step; b_lo += 4 * b_step; acc_first = 1; n -= 4; }
  if (n >= 2) { mma_burst<2>(d_tmem, a_lo, b_lo, hi, idesc, acc_first, a_step, b_step); a_lo += 2 * a_step; b_lo += 2 * b_step; acc_first = 1; n -= 2; }
  if (n >= 1) mma_burst<1>(d_tmem, a_lo, b_lo, hi, idesc, acc_first, a_step, b_step);
}

template <int NT>
__global__ void __launch_bounds__(192 + 160 * NT, 1) ffn_tc_kernel(FfnTcParams p, FfnTcGeom g) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = g.C, H = g.H, KT = g.KT, NC = g.NC, KS = g.KS, AR = g.AR, TS = g.TS, NS = g.NS;
  const int KH = g.KH, TPS = g.TPS;
  constexpr int NA = NT == 2 ? NT + TC_A_EXTRA : 2;    // A-tile slots (single-tile variant: plain double buffering)
  const uint32_t sbase = smem_u32(smem);
  float* tab_b1 = reinterpret_cast<float*>(smem + g.off_tab);
  float* tab_b2 = tab_b1 + 2 * H;
  float* tab_gamma = tab_b2 + C;
  const uint32_t bar0 = sbase + g.off_bar;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  // barrier map: 0..7 w_full, 8..15 w_empty, 16..19 a_full, 20..23 a_empty, 24..25 d1_full[tile], 26..27 d1_empty,
  // 28..29 g_full, 30..31 g_empty, 32..33 d2_full, 34..35 d2_empty; slot 48: TMEM base address
  const int W_FULL = 0, W_EMPTY = 8, A_FULL = 16, A_EMPTY = 20, D1_FULL = 24, D1_EMPTY = 26, G_FULL = 28, G_EMPTY = 30,
            D2_FULL = 32, D2_EMPTY = 34, F_DONE = 36;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + g.off_bar + 8 * 48);

  for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) tab_b1[i] = p.b1[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { tab_b2[i] = p.b2[i]; tab_gamma[i] = p.gamma[i]; }
  {  // hidden tiles: rows >= 128 are only ever read for discarded output rows; keep them finite
    uint32_t* gz = reinterpret_cast<uint32_t*>(smem + g.off_g);
    for (uint32_t i = threadIdx.x; i < NT * g.g_buf_bytes / 4; i += blockDim.x) gz[i] = 0u;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(BAR(W_FULL + i), 1); mbar_init(BAR(W_EMPTY + i), 2 * NT); }
    for (int i = 0; i < 4; ++i) { mbar_init(BAR(A_FULL + i), 128); mbar_init(BAR(A_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR(D1_FULL + i), 1); mbar_init(BAR(D1_EMPTY + i), 128);
      mbar_init(BAR(G_FULL + i), 128); mbar_init(BAR(G_EMPTY + i), 1);
      mbar_init(BAR(D2_FULL + i), 1); mbar_init(BAR(D2_EMPTY + i), 128); mbar_init(BAR(F_DONE + i), 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  unsigned long long* const tr = (blockIdx.x == 0 && lane == 0) ? g_trace : nullptr;   // diagnostic event trace
  const int n_pairs = (p.n_tiles + NT - 1) / NT;
  const int n_iter = (n_pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint32_t d2_col0 = 2u * 2 * TC_HC;  // after the two D1 tiles

  if (warp == 0) {
    // ===================== weight loader (whole warp runs the loop, one elected lane issues) =====================
    // stage order = MMA issue order: for the running chunk index q (continuous across the CTA's tile pairs, so the
    // first conv1d chunk of the next pair is issued BEFORE the last transposed-conv chunk of this one and the tensor
    // pipe never drains at a pair boundary): W1(q mod NC), then W2((q - 1) mod NC); W2(NC - 1) closes the stream.
    {
      uint32_t slot = 0, ph = 0;
      const int n1 = KT * KH;
      const char* const w2base = p.img + (size_t)NC * n1 * g.stage_bytes;
      const int Q = n_iter * NC;
      _Pragma("unroll 1") for (int q = 0; q <= Q; ++q) {
        _Pragma("unroll 1") for (int part = 0; part < 2; ++part) {
          const char* src; int n;
          if (part == 0) { if (q == Q) continue; src = p.img + (size_t)(q % NC) * n1 * g.stage_bytes; n = n1; }
          else { if (q == 0) continue; src = w2base + (size_t)((q - 1) % NC) * KS * g.stage_bytes; n = KS; }
          _Pragma("unroll 1") for (int s = 0; s < n; ++s, src += g.stage_bytes) {
            mbar_wait(BAR(W_EMPTY + slot), ph ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(BAR(W_FULL + slot), g.stage_bytes);
              bulk_g2s(sbase + g.off_w + slot * g.stage_bytes, src, g.stage_bytes, BAR(W_FULL + slot));
            }
            __syncwarp();
            if (++slot == (uint32_t)NS) { slot = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp <= NT + 1) {
    // ===================== MMA issuers: one warp per tile slot for the conv1d MMAs (M1), one for the transposed conv (M2) ====
    // Issuing a tcgen05.mma costs the issuing thread 80-150 clocks in this kernel (descriptor words travel through R2UR,
    // mbarrier polls and commits sit in between; profiles/mma_rate_bench.cu), more than the 64 clocks an
    // M128 x N128 x K16 MMA occupies the tensor pipe -- one issuing warp left the pipe idle half of the time.  Each
    // tile's accumulators, A / G tiles and barriers are private to the tile, and M1 / M2 of a tile use disjoint
    // barrier sets, so the issue work is spread over NT + 1 warps (16 warps in all: four per scheduler, which keeps
    // 128 registers per thread); they share only the weight ring, whose stages are consumed in image order.
    // ONE elected lane runs each warp's whole persistent loop (elect.sync once, at the top: ptxas then knows a single
    // thread is active and emits bare UTCHMMA / UTCBAR).
    if (elect_one()) {
      const bool is_m1 = warp <= NT;
      const int t = is_m1 ? warp - 1 : 0;
      const uint32_t idesc1 = instr_desc(128, 2 * TC_HC), idesc2 = instr_desc(128, C);
      // descriptor words: lo = (addr >> 4) | (LBO/16 << 16); hi = SBO/16 | version(1) << 14
      const uint32_t hi = (128u >> 4) | (1u << 14);
      const uint32_t lo_a = (uint32_t)AR << 16;            // LBO = AR*16 for A and G tiles
      const uint32_t lo_b1 = 128u << 16;                   // W1 stage: 128 rows -> LBO = 2048
      const uint32_t lo_b2 = (uint32_t)C << 16;            // W2 stage: C rows   -> LBO = C*16
      const uint32_t KK1 = C / KH / 16;                    // MMAs per W1 stage per tile
      const uint32_t w16 = (sbase + g.off_w) >> 4, stage16 = g.stage_bytes >> 4;
      const uint32_t gb = ((sbase + g.off_g) >> 4) + t * (g.g_buf_bytes >> 4);
      const uint32_t a16 = (sbase + g.off_a) >> 4, aslot16 = g.a_slot_bytes >> 4;
      const uint32_t tap16 = (uint32_t)(TC_HC / 8) * C;    // one W2 tap = 8 chunks * C rows * 16 B
      const uint32_t d1col = tmem + t * (2 * TC_HC), d2col = tmem + d2_col0 + t * C;
      unsigned long long* const mtr = (t == 0 && blockIdx.x == 0) ? g_trace : nullptr;   // the elected lanes trace
      uint32_t wslot = 0, wph = 0;                         // ring position of the NEXT stage in image order
      // Every stage is released by ALL MMA warps (W_EMPTY counts 2 * NT: NT arrivals from the warps of each kind): a
      // warp steps over a stage of the other kind by observing its full phase and arriving at once.  The ring then
      // moves in lock step with every warp, so a parity wait always refers to the phase right after the last one the
      // thread has seen (a warp that merely skipped slots could be lapped by the loader, or poll a slot two uses behind).
      auto skip = [&](int n) {
        for (int i = 0; i < n; ++i) {
          mbar_wait(BAR(W_FULL + wslot), wph);
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(BAR(W_EMPTY + wslot)), "r"(is_m1 ? 1 : NT) : "memory");
          if (++wslot == (uint32_t)NS) { wslot = 0; wph ^= 1; }
        }
      };
      const int n1 = KT * KH;
      if (is_m1) {
        uint32_t aslot = t, aph = 0;
        uint32_t qpar = 0;                                 // parity of the running chunk counter
        for (int it = 0; it < n_iter; ++it) {
          for (int c = 0; c < NC; ++c) {
            trace_event(mtr, 0, it * NC + c);
            if (c == 0) mbar_wait(BAR(A_FULL + aslot), aph);
            mbar_wait(BAR(D1_EMPTY + t), qpar ^ 1);
            tc_fence_after();
            trace_event(mtr, 1, it * NC + c);
            const uint32_t ab0 = a16 + aslot * aslot16;
            long long wacc1 = 0;
            for (int k = 0; k < KT; ++k)
              for (int hf = 0; hf < KH; ++hf) {
                const long long tw = mtr ? clock64() : 0;
                mbar_wait(BAR(W_FULL + wslot), wph);
                if (mtr) wacc1 += clock64() - tw;
                tc_fence_after();
                const uint32_t wb = w16 + wslot * stage16;
                mma_run((int)KK1, d1col, (ab0 + k + hf * KK1 * 2 * AR) | lo_a, wb | lo_b1, hi, idesc1, (uint32_t)(k | hf), 2u * AR,
                        256u);
                mma_commit(BAR(W_EMPTY + wslot));
                if (++wslot == (uint32_t)NS) { wslot = 0; wph ^= 1; }
              }
            mma_commit(BAR(D1_FULL + t));
            if (c == NC - 1) mma_commit(BAR(A_EMPTY + aslot));
            trace_event(mtr, 2, it * NC + c);
            if (mtr != nullptr && it * NC + c < 64) mtr[11 * 64 + it * NC + c] = (unsigned long long)wacc1;
            if (it + c > 0) skip(KS);                      // the M2 stages of the previous chunk (running index)
            qpar ^= 1;
          }
          for (int a = 0; a < NT; ++a)   // advance this tile's A slot by NT positions in the NA ring
            if (++aslot == (uint32_t)NA) { aslot = 0; aph ^= 1; }
        }
        if (n_iter > 0) skip(KS);                          // the M2 stages of the very last chunk
      } else {
        uint32_t qpar = 0;
        const int Q = n_iter * NC;
        int cc = NC - 1, it2 = -1;                         // chunk / tile pair of the transposed conv issued at step q
        for (int q = 0; q <= Q; ++q) {
          if (q < Q) skip(n1);                             // the M1 stages of chunk q
          if (q > 0) {
            trace_event(mtr, 7, it2 * NC + cc);
            _Pragma("unroll") for (int u = 0; u < NT; ++u) mbar_wait(BAR(G_FULL + u), qpar);
            trace_event(mtr, 8, it2 * NC + cc);
            if (cc == 0) _Pragma("unroll") for (int u = 0; u < NT; ++u) mbar_wait(BAR(D2_EMPTY + u), (uint32_t)((it2 & 1) ^ 1));
            tc_fence_after();
            long long wacc = 0;
            for (int s = 0; s < KS; ++s) {
              const long long tw = mtr ? clock64() : 0;
              mbar_wait(BAR(W_FULL + wslot), wph);
              if (mtr) wacc += clock64() - tw;
              tc_fence_after();
              const uint32_t wb = w16 + wslot * stage16;
              _Pragma("unroll") for (int u = 0; u < NT; ++u) {
                for (int tl = 0; tl < TPS; ++tl) {
                  const int tap = s * TPS + tl;
                  if (tap >= KT) break;
                  mma_run(TC_HC / 16, d2col + u * C, (gb + u * (g.g_buf_bytes >> 4) + tap) | lo_a, (wb + tl * tap16) | lo_b2, hi,
                          idesc2, (uint32_t)(cc | tap), 2u * AR, 2u * C);
                }
                mma_commit(BAR(W_EMPTY + wslot));          // one arrival per tile
              }
              if (++wslot == (uint32_t)NS) { wslot = 0; wph ^= 1; }
            }
            _Pragma("unroll") for (int u = 0; u < NT; ++u) mma_commit(BAR(G_EMPTY + u));
            if (cc == NC - 1) _Pragma("unroll") for (int u = 0; u < NT; ++u) mma_commit(BAR(D2_FULL + u));
            trace_event(mtr, 9, it2 * NC + cc);
            if (mtr != nullptr && it2 * NC + cc < 64) mtr[12 * 64 + it2 * NC + cc] = (unsigned long long)wacc;
            qpar ^= 1;
          }
          if (++cc == NC) { cc = 0; ++it2; }               // step q + 1 issues the transposed conv of chunk q
        }
      }
    }
  } else if (warp < 6 + NT) {
    // ===================== A producers / output writers =====================
    // x -> RMSGroupNorm -> bf16 chunk-major A tile for the NEXT tile pair, then the final epilogue of the current
    // pair (transposed-conv accumulator + bias + residual -> y), so the SwiGLU warps never leave the chunk loop.
    const int tp = threadIdx.x - 32 * (2 + NT);  // 0..127
    const int G = g.G, D = C / G;
    const float rs = rsqrtf((float)D);
    uint32_t slot = 0, ph = 0;
    // (code size matters: 16 warps run five different programs, and every unrolled copy of these bodies competes
    //  for the instruction cache -- ncu showed 'no instruction' as a top stall reason of the epilogue warps)
    constexpr int AHEAD = NA / NT;
    auto produce = [&](int it) {
      _Pragma("unroll 1") for (int t = 0; t < NT; ++t) {
        mbar_wait(BAR(A_EMPTY + slot), ph ^ 1);
        if (it >= AHEAD) mbar_wait(BAR(F_DONE + t), (uint32_t)(((it - AHEAD) & 1)));
        if (warp == NT + 2 && tr != nullptr && it < 8) tr[13 * 64 + 32 + 4 * it + 2 * t] = (unsigned long long)clock64();
        uint8_t* at = smem + g.off_a + (size_t)slot * g.a_slot_bytes;
        const long long tile = ((long long)blockIdx.x + (long long)it * gridDim.x) * NT + t;
        const long long r0 = tile * TS;
        const int s0 = (int)(r0 / p.P), j0 = (int)(r0 - (long long)s0 * p.P);   // one 64-bit division per tile
        auto locate = [&](int item, int& row, int& grp) -> const float* {        // source of a (row, group) item or nullptr
          row = item / G; grp = item - row * G;
          if (item >= AR * G || r0 + row >= p.R) return nullptr;
          int sq = s0, j = j0 + row;
          while (j >= p.P) { j -= p.P; ++sq; }
          if (j < KT - 1) return nullptr;                                          // zero rows between two sequences
          return p.x + p.map.base(sq) + (long long)(j - (KT - 1)) * p.map.pos_stride + grp * D;
        };
        if (D <= 32) {
          // A (row, group) item is <= 32 channels: fetched with back-to-back independent 128-bit loads, kept in
          // registers between the sum of squares and the scaling (x read once).  The loads of the NEXT item are in
          // flight while this one is normalised and stored: two memory latencies overlap instead of one per pass.
          auto fetch = [&](const float* src, float4 (&v)[8]) {
#pragma unroll
            for (int d = 0; d < 8; ++d)
              v[d] = (src != nullptr && 4 * d < D) ? __ldg(reinterpret_cast<const float4*>(src + 4 * d)) : make_float4(0.f, 0.f, 0.f, 0.f);
          };
          auto emit = [&](const float4 (&v)[8], int row, int grp) {
            float ss = 0.f;
#pragma unroll
            for (int d = 0; d < 8; ++d) ss += v[d].x * v[d].x + v[d].y * v[d].y + v[d].z * v[d].z + v[d].w * v[d].w;
            const float inv = 1.f / (sqrtf(ss) * rs + p.eps);
#pragma unroll
            for (int d = 0; d < 8; ++d) {
              if (4 * d < D) {
                const int c0 = grp * D + 4 * d;
                const float4 gm = *reinterpret_cast<const float4*>(tab_gamma + c0);
                uint2 pk;
                pk.x = pack_bf16(v[d].x * inv * gm.x, v[d].y * inv * gm.y);
                pk.y = pack_bf16(v[d].z * inv * gm.z, v[d].w * inv * gm.w);
                *reinterpret_cast<uint2*>(at + ((size_t)(c0 >> 3) * AR + row) * 16 + (c0 & 7) * 2) = pk;
              }
            }
          };
          float4 va[8], vb[8];
          int rowa, grpa, rowb, grpb;
          fetch(locate(tp, rowa, grpa), va);
          _Pragma("unroll 1") for (int item = tp; item < AR * G; item += 128) {
            fetch(locate(item + 128, rowb, grpb), vb);
            emit(va, rowa, grpa);
#pragma unroll
            for (int d = 0; d < 8; ++d) va[d] = vb[d];
            rowa = rowb; grpa = grpb;
          }
        } else {
          for (int item = tp; item < AR * G; item += 128) {
            int row, grp;
            const float* src = locate(item, row, grp);
            const bool valid = src != nullptr;
            float ss = 0.f;
            if (valid)
              for (int d = 0; d < D; d += 4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(src + d));
                ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
              }
            const float inv = 1.f / (sqrtf(ss) * rs + p.eps);
            for (int d = 0; d < D; d += 4) {
              uint2 pk = make_uint2(0u, 0u);
              const int c0 = grp * D + d;
              if (valid) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(src + d));  // L1 hit
                const float4 gm = *reinterpret_cast<const float4*>(tab_gamma + c0);
                pk.x = pack_bf16(v.x * inv * gm.x, v.y * inv * gm.y);
                pk.y = pack_bf16(v.z * inv * gm.z, v.w * inv * gm.w);
              }
              *reinterpret_cast<uint2*>(at + ((size_t)(c0 >> 3) * AR + row) * 16 + (c0 & 7) * 2) = pk;
            }
          }
        }
        if (warp == NT + 2 && tr != nullptr && it < 8) tr[13 * 64 + 32 + 4 * it + 2 * t + 1] = (unsigned long long)clock64();
        fence_proxy_async();
        mbar_arrive(BAR(A_FULL + slot));
        if (++slot == (uint32_t)NA) { slot = 0; ph ^= 1; }
      }
    };
    // The A tiles run AHEAD pairs ahead.  The final epilogue of pair `it` comes FIRST in every round: the transposed
    // conv of the next pair's first chunk waits for it (D2_EMPTY), about one chunk period after D2_FULL.
    // The A tiles run AHEAD pairs ahead; a slot is rewritten once the MMAs that read it are done (A_EMPTY) and the
    // epilogue group of that tile slot has finished its output pass, which borrows the slot as a staging strip (F_DONE).
    _Pragma("unroll 1") for (int it = 0; it < n_iter; ++it) produce(it);
  } else {
    // ===================== epilogue groups (one per tile slot) =====================
    const int t = (warp - (6 + NT)) >> 2;
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int m = quarter * 32 + lane;       // tile row
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    uint8_t* gt = smem + g.off_g + (size_t)t * g.g_buf_bytes;
    uint32_t qpar = 0;
    // ---- output pass of one finished tile: transposed-conv accumulator + bias + residual -> y ----
    // TMEM hands every thread one ROW of the tile; a warp instruction that touches 32 rows of x costs 32 L1 tag cycles.
    // So each warp passes its 32 rows through a padded strip in shared memory (the tile's own A slot: every MMA that
    // read it completed before D2_FULL; the producers wait for F_DONE before they rewrite it), 32 columns per step, and
    // walks the strip with 8 lanes per row (128 contiguous bytes): 4 lines per load / store instead of 32.  The
    // residual of the next step is requested before this step's stores.
    constexpr uint32_t FPITCH = 32 * 4 + 16;               // strip row: 32 fp32 + 16 B (conflict-free both ways)
    const int frow = lane >> 3, fcol = (lane & 7) * 4;      // coalesced phase: row within a group of 4, first column
    unsigned long long* const etr = warp == 6 + NT ? tr : nullptr;
    auto finish = [&](int it) {
      const long long tile = ((long long)blockIdx.x + (long long)it * gridDim.x) * NT + t;
      if ((C & 31) == 0 && g.a_slot_bytes >= 128 * FPITCH) {   // the strips of the four warps must fit in the A slot
        const int steps = C / 32;
        int pos[8];                                        // stream position (x / y offset = pos * C) of row quarter*32 + 4*i + frow
        _Pragma("unroll") for (int i = 0; i < 8; ++i) {
          const int mm = quarter * 32 + 4 * i + frow;
          const long long ro = tile * TS + mm;
          pos[i] = -1;
          if (mm < TS && ro < p.R) {
            const int sq = (int)(ro / p.P), j = (int)(ro - (long long)sq * p.P);
            if (j < p.S) pos[i] = (int)((p.map.base(sq) + (long long)j * p.map.pos_stride) / C);
          }
        }
        auto fetch = [&](float4 (&xv)[8], int st) {
          _Pragma("unroll") for (int i = 0; i < 8; ++i)
            xv[i] = pos[i] >= 0 ? __ldg(reinterpret_cast<const float4*>(p.x + (long long)pos[i] * C + st * 32 + fcol))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        uint8_t* strip = smem + g.off_a + (size_t)(((uint32_t)(it * NT + t)) % NA) * g.a_slot_bytes + (size_t)(quarter * 32) * FPITCH;
        float4 xa[8], xb[8];
        fetch(xa, 0);                                        // in flight while this warp waits for the accumulator
        mbar_wait(BAR(D2_FULL + t), (uint32_t)(it & 1));
        tc_fence_after();
        trace_event(etr, 13, it);
        _Pragma("unroll 1") for (int st = 0; st < steps; ++st) {
          {
            uint32_t r[32];
            tmem_ld32(lane_addr + d2_col0 + t * C + st * 32, r);
            tc_wait_ld();
            _Pragma("unroll") for (int e = 0; e < 8; ++e)
              *reinterpret_cast<uint4*>(strip + (size_t)lane * FPITCH + e * 16) = make_uint4(r[4 * e], r[4 * e + 1], r[4 * e + 2], r[4 * e + 3]);
          }
          if (st == steps - 1) {                             // D2 of this tile fully read
            tc_fence_before();
            mbar_arrive(BAR(D2_EMPTY + t));
          }
          __syncwarp();
          if (st + 1 < steps) fetch(xb, st + 1);
          const float4 b = *reinterpret_cast<const float4*>(tab_b2 + st * 32 + fcol);
          _Pragma("unroll") for (int i = 0; i < 8; ++i) {
            const float4 d = *reinterpret_cast<const float4*>(strip + (size_t)(4 * i + frow) * FPITCH + fcol * 4);
            if (pos[i] >= 0) {
              float4 o = xa[i];
              o.x += d.x + b.x; o.y += d.y + b.y; o.z += d.z + b.z; o.w += d.w + b.w;
              *reinterpret_cast<float4*>(p.y + (long long)pos[i] * C + st * 32 + fcol) = o;
            }
          }
          __syncwarp();
          _Pragma("unroll") for (int i = 0; i < 8; ++i) xa[i] = xb[i];
        }
      } else {
        mbar_wait(BAR(D2_FULL + t), (uint32_t)(it & 1));
        tc_fence_after();
        const long long ro = tile * TS + m;
        float* dst = nullptr;
        const float* res = nullptr;
        if (m < TS && ro < p.R) {
          const int sq = (int)(ro / p.P), i = (int)(ro - (long long)sq * p.P);
          if (i < p.S) {
            const long long off = p.map.base(sq) + (long long)i * p.map.pos_stride;
            dst = p.y + off; res = p.x + off;
          }
        }
        _Pragma("unroll 1") for (int c0 = 0; c0 < C; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(lane_addr + d2_col0 + t * C + c0, r);
          tc_wait_ld();
          if (dst != nullptr) {
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              float4 xv = __ldg(reinterpret_cast<const float4*>(res + c0 + e));
              xv.x += __uint_as_float(r[e]) + tab_b2[c0 + e];
              xv.y += __uint_as_float(r[e + 1]) + tab_b2[c0 + e + 1];
              xv.z += __uint_as_float(r[e + 2]) + tab_b2[c0 + e + 2];
              xv.w += __uint_as_float(r[e + 3]) + tab_b2[c0 + e + 3];
              *reinterpret_cast<float4*>(dst + c0 + e) = xv;
            }
          }
        }
        tc_fence_before();
        mbar_arrive(BAR(D2_EMPTY + t));
      }
      trace_event(etr, 14, it);
      mbar_arrive(BAR(F_DONE + t));                          // the strip (this tile's A slot) may be rewritten
    };
    // Running chunk index q across the CTA's tile pairs.  The output pass of pair it - 1 runs right after the SwiGLU of
    // chunk 0 of pair it: the MMA warps already have D1 of that chunk back and the hidden tile of chunk 0, so the
    // tensor pipe keeps working on the next pair while this group writes the previous one out.
    const int Q = n_iter * NC;
    int c = 0, it = 0;
    _Pragma("unroll 1") for (int q = 0; q <= Q; ++q) {
      if (q < Q) {
        mbar_wait(BAR(D1_FULL + t), qpar);
        tc_fence_after();
        trace_event(etr, 3, it * NC + c);
        uint32_t packed[TC_HC / 2];
        const float* bv = tab_b1 + c * TC_HC;
        const float* bg = tab_b1 + H + c * TC_HC;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t rv[32], rg[32];
          tmem_ld32(lane_addr + t * (2 * TC_HC) + half * 32, rv);
          tmem_ld32(lane_addr + t * (2 * TC_HC) + TC_HC + half * 32, rg);
          tc_wait_ld();
          if (half == 1) {  // D1 fully read: hand it back to the MMA thread before the math
            tc_fence_before();
            mbar_arrive(BAR(D1_EMPTY + t));
            trace_event(etr, 4, it * NC + c);
          }
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float hv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const float val = __uint_as_float(rv[i + u]) + bv[half * 32 + i + u];
              const float gate = __uint_as_float(rg[i + u]) + bg[half * 32 + i + u];
              hv[u] = swiglu_fast(val, gate);                         // value * SiLU(gate), :648-649
            }
            packed[half * 16 + (i >> 1)] = pack_bf16(hv[0], hv[1]);
          }
        }
        trace_event(etr, 10, it * NC + c);
        mbar_wait(BAR(G_EMPTY + t), qpar ^ 1);   // transposed-conv MMAs of the previous chunk are done with G
        trace_event(etr, 5, it * NC + c);
#pragma unroll
        for (int ch = 0; ch < TC_HC / 8; ++ch)
          *reinterpret_cast<uint4*>(gt + ((size_t)ch * AR + m) * 16) =
              make_uint4(packed[ch * 4], packed[ch * 4 + 1], packed[ch * 4 + 2], packed[ch * 4 + 3]);
        fence_proxy_async();
        mbar_arrive(BAR(G_FULL + t));
        trace_event(etr, 6, it * NC + c);
        qpar ^= 1;
      }
      if (q > 0 && c == 0) finish(it - 1);
      if (++c == NC) { c = 0; ++it; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// --------------------------------------------------------------------------------------------
// Self-test of the tcgen05 plumbing (descriptors, row-shifted taps, bulk copy, TMEM round trip):
// D[128, N] = sum_tap A[m + tap, :] . B_tap[n, :]   (mode 0: B K-major, staged by a bulk copy)
// D[128, N] = A[m, :] . V[:, n]                     (mode 1: V MN-major, rows of V are the K index)
// same with A staged in TMEM by tcgen05.st          (mode 2: the P.V form of attn_tc2_kernel; N <= 128)
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                             const __nv_bfloat16* __restrict__ Bimg,
                                                             float* __restrict__ D, int N, int Kd, int taps, int mode) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int AR = 128 + taps - 1;
  const uint32_t a_bytes = (uint32_t)(Kd / 8) * AR * 16;
  const uint32_t b_rows = mode == 0 ? N : Kd;
  const uint32_t b_cols = mode == 0 ? Kd : N;
  const uint32_t b_tap_bytes = (b_cols / 8) * b_rows * 16;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((a_bytes + 127) & ~127u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb + ((taps * b_tap_bytes + 127) & ~127u));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  for (int i = threadIdx.x; i < AR * Kd; i += blockDim.x) {
    const int r = i / Kd, c = i % Kd;
    *reinterpret_cast<__nv_bfloat16*>(sa + ((size_t)(c >> 3) * AR + r) * 16 + (c & 7) * 2) = __float2bfloat16_rn(A[i]);
  }
  if (mode >= 1) {
    for (int i = threadIdx.x; i < Kd * N; i += blockDim.x) {
      const int r = i / N, c = i % N;  // V[r = k index][c = n index]
      *reinterpret_cast<__nv_bfloat16*>(sb + ((size_t)(c >> 3) * Kd + r) * 16 + (c & 7) * 2) = __float2bfloat16_rn(B[i]);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (mode == 2) {
    // row m of A as packed bf16 pairs in TMEM columns [256, 256 + Kd/2): thread m owns lane m
    const int m = warp * 32 + lane;
    for (int c0 = 0; c0 < Kd; c0 += 32) {
      uint32_t r[16];
      for (int e = 0; e < 16; ++e)
        r[e] = c0 + 2 * e < Kd ? pack_bf16(A[(size_t)m * Kd + c0 + 2 * e], A[(size_t)m * Kd + c0 + 2 * e + 1]) : 0u;
      tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + 256 + c0 / 2, r);
    }
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (threadIdx.x == 0) {
    if (mode == 0) {
      mbar_arrive_expect_tx(smem_u32(&bars[0]), taps * b_tap_bytes);
      bulk_g2s(smem_u32(sb), Bimg, taps * b_tap_bytes, smem_u32(&bars[0]));
      mbar_wait(smem_u32(&bars[0]), 0);
      tc_fence_after();
      const uint32_t idesc = instr_desc(128, N);
      for (int t = 0; t < taps; ++t)
        for (int kk = 0; kk < Kd / 16; ++kk) {
          const uint64_t ad = smem_desc(smem_u32(sa) + t * 16 + kk * 2 * AR * 16, AR * 16, 128);
          const uint64_t bd = smem_desc(smem_u32(sb) + t * b_tap_bytes + kk * 2 * N * 16, N * 16, 128);
          mma_ss(tmem, ad, bd, idesc, (t | kk) != 0);
        }
    } else {
      const uint32_t idesc = instr_desc(128, N, /*b_mn_major=*/true);
      for (int kk = 0; kk < Kd / 16; ++kk) {
        const uint64_t ad = smem_desc(smem_u32(sa) + kk * 2 * AR * 16, AR * 16, 128);
        // MN-major: LBO = 128 B between 8-row K groups, SBO = Kd*16 between 8-column N groups
        const uint64_t bd = smem_desc(smem_u32(sb) + kk * 16 * 16, 128, Kd * 16);
        if (mode == 1) mma_ss(tmem, ad, bd, idesc, kk != 0);
        else mma_ts(tmem, tmem + 256 + kk * 8, bd, idesc, kk != 0);
      }
    }
    mma_commit(smem_u32(&bars[1]));
  }
  mbar_wait(smem_u32(&bars[1]), 0);
  tc_fence_after();
  const int m = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tc_wait_ld();
    for (int e = 0; e < 16; ++e) D[(size_t)m * N + c0 + e] = __uint_as_float(r[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

__global__ void tc_selftest_pack_kernel(const float* __restrict__ B, __nv_bfloat16* __restrict__ img, int N, int Kd, int taps) {
  // B [taps][N][Kd] fp32 -> per tap chunk-major [Kd/8][N][8] bf16
  const int total = taps * N * Kd;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int t = i / (N * Kd), n = (i / Kd) % N, c = i % Kd;
    img[(size_t)t * N * Kd + ((size_t)(c >> 3) * N + n) * 8 + (c & 7)] = __float2bfloat16_rn(B[i]);
  }
}

inline int tc_selftest(const float* A, const float* B, float* D, void* scratch, int N, int Kd, int taps, int mode,
                       cudaStream_t st) {
  TFL_CHECK(N % 16 == 0 && N <= 256 && Kd % 16 == 0 && taps >= 1 && taps <= 8, "selftest shape");
  TFL_CHECK(mode == 0 || taps == 1, "modes 1 and 2 use a single tap");
  TFL_CHECK(mode != 2 || (N <= 128 && Kd <= 256), "mode 2 shape");
  const int AR = 128 + taps - 1;
  const size_t a_bytes = ((size_t)(Kd / 8) * AR * 16 + 127) & ~(size_t)127;
  const size_t b_bytes = ((size_t)taps * N * Kd * 2 + 127) & ~(size_t)127;
  const size_t smem = a_bytes + b_bytes + 256;
  TFL_CHECK(smem <= (size_t)TC_SMEM_MAX, "selftest does not fit in shared memory");
  TFL_CUDA(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (mode == 0) tc_selftest_pack_kernel<<<64, 256, 0, st>>>(B, (__nv_bfloat16*)scratch, N, Kd, taps);
  tc_selftest_kernel<<<1, 128, smem, st>>>(A, B, (const __nv_bfloat16*)scratch, D, N, Kd, taps, mode);
  TFL_LAUNCH_CHECK();
  return 0;
}

// --------------------------------------------------------------------------------------------
inline size_t tc_workspace_bytes(const tfl_plan* pl, int B, int Tf, int F) {
  return (size_t)B * Tf * F * pl->cfg.emb_dim * sizeof(float);  // second residual buffer (ping-pong)
}

int tfl_option(int key);
inline int tc_ffn2_dispatch(const tfl_plan* pl, const FfnTcParams& p, int C, int H, int K, int G, cudaStream_t st);

// y = x + ConvSwiGLU(RMSGroupNorm(x)); x and y must be distinct buffers.
inline int tc_ffn(const tfl_plan* pl, const char* packed, int layer, int axis, int j, const float* x, float* y, int B,
                  int Tf, int F, cudaStream_t st) {
  nvtxRangePushA("tfl::conv_swiglu_ffn[bf16]");
  struct Pop { ~Pop() { nvtxRangePop(); } } nvtx_pop;
  TFL_CHECK(x != y, "tc_ffn needs distinct input and output buffers");
  const tfl_config& c = pl->cfg;
  const FfnPack& f = pl->lay.paths[(size_t)layer * 2 + axis].ffn[j];
  FfnTcGeom g;
  TFL_CHECK(ffn_tc_geometry(c.emb_dim, f.hidden, c.conv_kernel, c.num_groups, &g),
            "bf16 tcgen05 FFN needs emb_dim %% 16 == 0 (<= 256), ffn_hidden %% 64 == 0, conv1d_kernel <= 8 "
            "(got emb_dim %d, hidden %d, kernel %d); use precision fp32 for this configuration",
            c.emb_dim, f.hidden, c.conv_kernel);
  const int S = axis == TFL_AXIS_FREQ ? F : Tf;
  const int nseq = axis == TFL_AXIS_FREQ ? B * Tf : B * F;
  FfnTcParams p;
  p.x = x; p.y = y; p.map = make_seq_map(axis, Tf, F, c.emb_dim);
  p.S = S; p.P = S + c.conv_kernel - 1; p.R = (long long)nseq * p.P;
  p.n_tiles = (int)((p.R + g.TS - 1) / g.TS);
  p.gamma = (const float*)(packed + f.gamma);
  p.b1 = (const float*)(packed + f.b1raw); p.b2 = (const float*)(packed + f.b2);
  p.img = packed + f.tc;
  p.eps = c.eps;
  if (f.tc2_ok && tfl_option(1 /* TFL_OPT_FFN_KERNEL */) == 2) {
    p.img = packed + f.tc2;
    return tc_ffn2_dispatch(pl, p, c.emb_dim, f.hidden, c.conv_kernel, c.num_groups, st);
  }
  if (g.NT == 2) TFL_CUDA(opt_in_smem(ffn_tc_kernel<2>, g.smem_bytes));
  else TFL_CUDA(opt_in_smem(ffn_tc_kernel<1>, g.smem_bytes));
  const int n_pairs = (p.n_tiles + g.NT - 1) / g.NT;
  const int grid = n_pairs < pl->sm_count ? n_pairs : pl->sm_count;
  if (g.NT == 2) ffn_tc_kernel<2><<<grid, g.threads, g.smem_bytes, st>>>(p, g);
  else ffn_tc_kernel<1><<<grid, g.threads, g.smem_bytes, st>>>(p, g);
  TFL_LAUNCH_CHECK();
  return 0;
}

}  // namespace tfl
