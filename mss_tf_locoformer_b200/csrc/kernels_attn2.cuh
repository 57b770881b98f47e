// K5 (v2)  attn_tc2_kernel: softmax(q k^T) v per (sequence, head) on tcgen05 / TMEM
// (models/mss_tflocoformer.py:523-531), same bf16 tile images in / out as kernels_attn.cuh.
//
// With head_dim 32 there are only 128 MMA FLOPs per exponential, so the kernel is paced by the 16/clk/SM exp2 unit
// and by whatever else the softmax threads must issue per score.  v2 strips that to the minimum:
//   * P never touches shared memory: the softmax threads write bf16 probabilities straight into TMEM
//     (tcgen05.st) and P.V runs with its A operand in TMEM -- no st.shared, no proxy fence, and the shared-memory
//     pipe only serves Q / K / V operand reads;
//   * O accumulates in TMEM over the whole key range (accumulating MMAs); nothing is folded per unit.  The softmax
//     reference m_ref is the row maximum of the first 32 keys and is only raised -- with a rescale of the TMEM
//     accumulator by the softmax thread itself -- when a later score exceeds it by more than 2^ATT2_TH
//     (bf16 / fp32 share an 8-bit exponent: a stale reference costs range, not precision);
//   * the work unit is a HALF of 32 keys (S = M128 x N32, one tcgen05.ld, one max3 tree, 32 exp2, tcgen05.st), so a
//     thread holds 32 + 8 registers of tile data and FOUR softmax groups (16 warps, 4 per SM sub-partition) fit --
//     enough warps in flight to keep the exp2 unit busy while others wait on TMEM loads or barriers;
//   * S, P share three rotating 32-column TMEM buffers per group: P_k overwrites the first 16 columns of the buffer
//     S_k was read from, and S_{k+3} is issued into it right behind P_k V_k (MMAs execute in issue order).  Scores
//     are therefore produced three halves ahead of their use: no softmax warp ever waits for the tensor pipe, and
//     the four warps of a group need not run in lock step;
//   * the key range ends in a half of N = ceil16(remaining keys) columns (1025 = 32*32 + 1, 259 = 8*32 + 3).
//
// One persistent CTA per SM; work item = (sequence, group of HG heads, group of QG query tiles), QG * HG = 4:
// the frequency axis (8 query tiles) runs 4 tiles of one head per item, the time axis (2 tiles) 2 tiles x 2 heads.
//   warps 0-15    four softmax groups, one query row per thread
//   warps 16-17   one MMA warp per PAIR of groups: S = Q K^T (64 keys ahead of the softmax), O += P_half V
//   warp 18       loader: Q tiles (double-buffered per item), K/V ring (one stage = 128 keys of K and V per head)
// (19 warps: at most 5 per SM sub-partition, which is what leaves 96 registers per thread -- with 21 warps the
// softmax loop spilled its running sums to local memory, 7 % of all instructions issued.)
// TMEM per group g (128 columns): three S/P buffers [0,32) [32,64) [64,96), O [96,96+HDP)
#pragma once
#include "kernels_attn.cuh"

namespace tfl {

struct Attn2Params {
  const __nv_bfloat16* qkv; __nv_bfloat16* o;
  int nseq, heads, L, NTL, HDP;
  int NQT;       // full query tiles handled here (trailing rows may go to attn_tail_rows_kernel)
  int QG, HG;    // query tiles / heads per work item
  int NQG, NHG;  // query-tile groups / head groups per sequence
  int NS;        // K/V ring stages
  int n_items;
};

#ifndef ATT2_PREFETCH
#define ATT2_PREFETCH 0   // 1: read the scores of half k + 1 back before the exp2 pass of half k (measured SLOWER, r02)
#endif
#ifndef ATT2_POLY_MASK
#define ATT2_POLY_MASK 0x00  // bit i: pair i of every 8 pairs of a half takes its two exponentials from the FMA pipe (exp2_poly) instead of MUFU.EX2
#endif
#ifndef ATT2_QUARTER
#define ATT2_QUARTER 0    // 1: 16-key quarters, the read-back of the next quarter in flight during the exp2 pass of this one (measured SLOWER, r02: 4.66 vs 4.47 ms per frequency-axis sub-block)
#endif
constexpr int ATT2_G = 4;
constexpr int ATT2_MMA_WARPS = ATT2_G / 2;                           // one MMA warp per pair of groups
constexpr int ATT2_THREADS = 32 * (4 * ATT2_G + 4);   // softmax warps; one warpgroup of MMA warps, loader and an idle warp
// Register split (setmaxnreg works on warpgroups of 4 consecutive warps and only re-splits the CTA's launch allocation
// of 640 x 96 registers): 4 * 104 + 64 = 480 = 5 * 96.
constexpr int ATT2_REGS_SOFTMAX = 104, ATT2_REGS_OTHER = 64;
static_assert(4 * ATT2_REGS_SOFTMAX + ATT2_REGS_OTHER <= 480, "register budget of the CTA");
#ifdef TFL_NO_SETMAXNREG   // bisecting aid: every warp keeps its launch allocation
#define ATT2_SETMAXNREG(dir, n) do { } while (0)
#else
#define ATT2_SETMAXNREG(dir, n) asm volatile("setmaxnreg." dir ".sync.aligned.u32 %0;" ::"n"(n))
#endif
constexpr float ATT2_TH = 16.f;    // raise the softmax reference when a score exceeds it by more than this (log2 units)

__device__ __forceinline__ float max3f(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// D[tmem] (+)= A[tmem] * B[smem] from 32-bit descriptor halves
__device__ __forceinline__ void mma_ts_lohi(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 bd, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], bd, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

inline uint32_t attn2_smem(int HDP, int HG, int* NS_out) {
  const uint32_t tile = (uint32_t)HDP * 128 * 2;
  const uint32_t fixed = 2u * ATT2_G * tile + 1024;
  const uint32_t stage = (uint32_t)HG * 2 * tile;
  int NS = (int)((TC_SMEM_MAX - fixed) / stage);
  if (NS > 8) NS = 8;
  *NS_out = NS;
  return fixed + (uint32_t)NS * stage;
}

// 2^x on the FMA pipe for x <= ATT2_TH (Cody-Waite: n = round(x) through the 1.5 * 2^23 magic add, f = x - n in
// [-0.5, 0.5], degree-3 minimax polynomial of 2^f with 7.5e-5 relative error -- 26 times below the bf16 rounding of the
// probability -- and n added straight into the exponent field).  The clamp keeps the exponent field positive: masked
// scores (-inf) come out as 2^-125 instead of 0, against key / value rows that are exact zeros.
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float xr = x + 12582912.f;
  const float f = x - (xr - 12582912.f);
  float p = fmaf(0.0551716648f, f, 0.2426111251f);
  p = fmaf(p, f, 0.6932609677f);
  p = fmaf(p, f, 0.9999280572f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(xr) << 23));
}

// One 32-key half for one query row: scores s[] (already read from TMEM buffer bb) -> max -> [rare: raise m_ref, rescale
// O] -> exp2 -> bf16 P written over the first 16 columns of the same buffer.  FULL: all 32 columns are keys; otherwise
// nkh are.  The read-back itself is issued by the caller one half AHEAD (ATT2_PREFETCH): TMEM reads run at 64 B/clk per
// SM -- 1024 clk for the 16 softmax warps of a half-step, the same as their 16 x 32 exp2 on the MUFU pipe -- so the
// two have to overlap inside every warp, not only across warps.
template <bool FULL>
__device__ __forceinline__ void attn2_half(uint32_t (&s)[32], uint32_t tcol, uint32_t bgrp, int bb, uint32_t ph, uint32_t kk,
                                           int nkh, bool first, int HDP, float& m_ref, float& l0, float& l1) {
  using namespace tc;
  constexpr uint32_t PV_DONE = 48;                               // byte offset inside the group's barrier block
  if (!FULL) {
#pragma unroll
    for (int i = 0; i < 32; ++i) if (i >= nkh) s[i] = 0xff800000u;   // -inf: columns beyond the sequence
  }
  float mx[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float t = max3f(__uint_as_float(s[8 * c]), __uint_as_float(s[8 * c + 1]), __uint_as_float(s[8 * c + 2]));
    t = max3f(t, __uint_as_float(s[8 * c + 3]), __uint_as_float(s[8 * c + 4]));
    mx[c] = max3f(t, __uint_as_float(s[8 * c + 5]), __uint_as_float(s[8 * c + 6]));
  }
  const float mxh = max3f(max3f(mx[0], mx[1], mx[2]), max3f(mx[3], __uint_as_float(s[7]), __uint_as_float(s[15])),
                          fmaxf(__uint_as_float(s[23]), __uint_as_float(s[31])));
  if (first) {
    m_ref = mxh;
  } else if (__any_sync(0xffffffffu, mxh > m_ref + ATT2_TH)) {
    // rare: raise the reference and rescale what has been accumulated so far.  Every P.V issued so far must have
    // landed; MMAs complete in issue order, so it is enough to wait for the one of the previous half.
    const float m_new = mxh > m_ref + ATT2_TH ? mxh : m_ref;
    const float alpha = fast_exp2(m_ref - m_new);
    const int pb = bb == 0 ? 2 : bb - 1;
    mbar_wait(bgrp + PV_DONE + 8 * pb, bb == 0 ? ph ^ 1 : ph);
    tc_fence_after();
    for (int c0 = 0; c0 < HDP; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tcol + 96 + c0, r);
      tc_wait_ld();                                              // (also drains a prefetched score read: harmless)
#pragma unroll
      for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
      tmem_st16(tcol + 96 + c0, r);
    }
    tc_wait_st();
    l0 *= alpha; l1 *= alpha;
    m_ref = m_new;
  }
  // Phase bookkeeping only (S of this half was issued after P.V of half kk-3, so that one has long completed): a
  // parity wait is only meaningful for the phase right after the last one this thread has observed.
  if (kk >= 3) mbar_wait(bgrp + PV_DONE + 8 * bb, ph ^ 1);
#pragma unroll
  for (int c = 0; c < 2; ++c) {                                  // 16 keys = one K step of the P.V MMA
    if (FULL || c == 0 || nkh > 16) {
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x0 = __uint_as_float(s[16 * c + 2 * i]) - m_ref, x1 = __uint_as_float(s[16 * c + 2 * i + 1]) - m_ref;
        const bool poly = ((ATT2_POLY_MASK >> i) & 1) != 0;
        const float p0 = poly ? exp2_poly(x0) : fast_exp2(x0);
        const float p1 = poly ? exp2_poly(x1) : fast_exp2(x1);
        l0 += p0; l1 += p1;
        pk[i] = pack_bf16(p0, p1);
      }
      tmem_st8(tcol + bb * 32 + c * 8, pk);
    }
  }
  tc_wait_st();
}

// ---- 16-key quarters (ATT2_QUARTER) ----
// TMEM read-back (64 B/clk per SM) and the exp2 unit each need ~1 024 clk per 32-key half-step of the 16 softmax warps,
// and a warp used them strictly one after the other (tcgen05.ld of 32 columns -> wait -> 32 exp2).  Here a half is
// read back as two 16-column quarters into the SAME 32 registers: while the exp2 pass of one quarter runs, the
// read-back of the next one (the second quarter of this half, or the first quarter of the next half) is in flight.
// No second score set (the 32 + 32 register variant, ATT2_PREFETCH, spilled) and no extra barrier in front of the
// exp2 burst.  The lazy reference makes this possible: a quarter's probabilities never wait for the maximum of the
// other quarter -- only the rare raise (a score more than 2^ATT2_TH above the reference) has to redo the first
// quarter's probabilities when the second quarter triggers it.
__device__ __forceinline__ float attn2_max16(const uint32_t (&s)[16]) {
  float m0 = max3f(__uint_as_float(s[0]), __uint_as_float(s[1]), __uint_as_float(s[2]));
  float m1 = max3f(__uint_as_float(s[3]), __uint_as_float(s[4]), __uint_as_float(s[5]));
  float m2 = max3f(__uint_as_float(s[6]), __uint_as_float(s[7]), __uint_as_float(s[8]));
  float m3 = max3f(__uint_as_float(s[9]), __uint_as_float(s[10]), __uint_as_float(s[11]));
  m0 = max3f(m0, __uint_as_float(s[12]), __uint_as_float(s[13]));
  m1 = max3f(m1, __uint_as_float(s[14]), __uint_as_float(s[15]));
  return max3f(max3f(m0, m1, m2), m3, m3);
}
template <bool SUM>
__device__ __forceinline__ void attn2_exp16(const uint32_t (&s)[16], float m_ref, uint32_t (&pk)[8], float& l0, float& l1) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float p0 = fast_exp2(__uint_as_float(s[2 * i]) - m_ref);
    const float p1 = fast_exp2(__uint_as_float(s[2 * i + 1]) - m_ref);
    if (SUM) { l0 += p0; l1 += p1; }
    pk[i] = tc::pack_bf16(p0, p1);
  }
}
// rare: raise the reference to m_new and rescale what has been accumulated (O in TMEM when have_o, the row sums)
__device__ __forceinline__ void attn2_raise(uint32_t tcol, uint32_t bgrp, int bb, uint32_t ph, bool have_o, int HDP,
                                            float m_new, float& m_ref, float& l0, float& l1) {
  using namespace tc;
  constexpr uint32_t PV_DONE = 48;
  const float alpha = fast_exp2(m_ref - m_new);
  if (have_o) {
    // every P.V issued so far must have landed; MMAs complete in issue order: wait for the one of the previous half
    const int pb = bb == 0 ? 2 : bb - 1;
    mbar_wait(bgrp + PV_DONE + 8 * pb, bb == 0 ? ph ^ 1 : ph);
    tc_fence_after();
    for (int c0 = 0; c0 < HDP; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tcol + 96 + c0, r);
      tc_wait_ld();                                              // (also drains a score read-back in flight: harmless)
#pragma unroll
      for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
      tmem_st16(tcol + 96 + c0, r);
    }
    tc_wait_st();
  }
  l0 *= alpha; l1 *= alpha;
  m_ref = m_new;
}

template <int KS>   // K steps of S = Q K^T (head_dim padded to 16 * KS)
__global__ void __launch_bounds__(ATT2_THREADS, 1) attn_tc2_kernel(Attn2Params p) {
  using namespace tc;
  constexpr int G = ATT2_G;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HDP = p.HDP, NTL = p.NTL, QG = p.QG, HG = p.HG, NS = p.NS;
  const int NH = (p.L + 31) / 32;                                // 32-key halves per sequence
  const int NST = NTL;                                           // K/V ring stages (128 keys) per item
  const int n_last = p.L - (NH - 1) * 32;                        // keys in the last half (1..32)
  const uint32_t tile_bytes = (uint32_t)HDP * 128 * 2;          // one Q / K / V tile (128 rows)
  const uint32_t stage_bytes = (uint32_t)HG * 2 * tile_bytes;   // per head: K tile, V tile
  const uint32_t off_q = 0, off_kv = 2u * G * tile_bytes, off_bar = off_kv + (uint32_t)NS * stage_bytes;
  const uint32_t sbase = smem_u32(smem);
  auto BAR = [&](int i) { return sbase + off_bar + 8u * i; };
  // barrier slots: ring 0..15, Q 16..19, then one block of 16 per group (a group's barriers are base + constant):
  //   +0..2 S_FULL[buffer]  +3..5 P_FULL[buffer]  +6..8 PV_DONE[buffer]
  const int KV_FULL = 0, KV_EMPTY = 8, Q_FULL = 16, Q_EMPTY = 18, GRP = 20;
  constexpr uint32_t S_FULL = 0, P_FULL = 24, PV_DONE = 48;      // byte offsets inside a group block
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + off_bar + 8 * (GRP + 16 * G));
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(BAR(KV_FULL + i), 1); mbar_init(BAR(KV_EMPTY + i), ATT2_MMA_WARPS); }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(Q_FULL + i), 1); mbar_init(BAR(Q_EMPTY + i), ATT2_MMA_WARPS); }
    for (int g = 0; g < G; ++g) {
      const uint32_t bg = BAR(GRP + 16 * g);
      for (int i = 0; i < 3; ++i) { mbar_init(bg + S_FULL + 8 * i, 1); mbar_init(bg + P_FULL + 8 * i, 128); mbar_init(bg + PV_DONE + 8 * i, 1); }
    }
    fence_barrier_init();
  }
  if (warp == 4 * G) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  griddep_wait();                                          // (PDL) the prologue overlapped the previous kernel's tail
  const size_t which_stride = (size_t)p.nseq * p.heads * NTL * HDP * 128;   // elements between the q, k, v planes

  // item -> (sequence, first head, first query tile); group g -> (head, tile).  Groups 2k and 2k+1 form a PAIR served
  // by one MMA warp; a pair runs the whole protocol when its first group has work (a second group without work --
  // ragged tile / head counts -- computes on whatever its buffers hold and stores nothing).
  auto decode = [&](int item, int& seq, int& h0, int& q0) {
    const int qg = item % p.NQG, t = item / p.NQG;
    q0 = qg * QG; h0 = (t % p.NHG) * HG; seq = t / p.NHG;
  };
  auto group_active = [&](int g, int h0, int q0) { return h0 + g / QG < p.heads && q0 + g % QG < p.NQT; };

  // (each role branch starts with its own setmaxnreg: ptxas bounds a region by the setmaxnreg that DOMINATES it)
  if (warp == 4 * G + ATT2_MMA_WARPS + 1) {
    // idle: only completes the warpgroup that gives its registers to the softmax groups
    ATT2_SETMAXNREG("dec", ATT2_REGS_OTHER);
  } else if (warp == 4 * G + ATT2_MMA_WARPS) {
    // ===================== loader =====================
    ATT2_SETMAXNREG("dec", ATT2_REGS_OTHER);
    uint32_t kslot = 0, kph = 0, qph = 0;
    int n_local = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++n_local) {
      int seq, h0, q0;
      decode(item, seq, h0, q0);
      const int b = n_local & 1;
      mbar_wait(BAR(Q_EMPTY + b), ((qph >> b) & 1) ^ 1);
      qph ^= 1u << b;
      if (elect_one()) {
        int n_act = 0;
        for (int g = 0; g < G; ++g) n_act += group_active(g, h0, q0) ? 1 : 0;
        mbar_arrive_expect_tx(BAR(Q_FULL + b), n_act * tile_bytes);
        for (int g = 0; g < G; ++g) {
          if (!group_active(g, h0, q0)) continue;
          const size_t sh = (size_t)seq * p.heads + h0 + g / QG;
          bulk_g2s(sbase + off_q + (b * G + g) * tile_bytes, p.qkv + (sh * NTL + q0 + g % QG) * HDP * 128, tile_bytes,
                   BAR(Q_FULL + b));
        }
      }
      __syncwarp();
      const int nh = min(HG, p.heads - h0);
      for (int j = 0; j < NST; ++j) {
        mbar_wait(BAR(KV_EMPTY + kslot), kph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(BAR(KV_FULL + kslot), nh * 2 * tile_bytes);
          for (int hh = 0; hh < nh; ++hh) {
            const __nv_bfloat16* src = p.qkv + (((size_t)seq * p.heads + h0 + hh) * NTL + j) * HDP * 128;
            const uint32_t dst = sbase + off_kv + kslot * stage_bytes + hh * 2 * tile_bytes;
            bulk_g2s(dst, src + which_stride, tile_bytes, BAR(KV_FULL + kslot));
            bulk_g2s(dst + tile_bytes, src + 2 * which_stride, tile_bytes, BAR(KV_FULL + kslot));
          }
        }
        __syncwarp();
        if (++kslot == (uint32_t)NS) { kslot = 0; kph ^= 1; }
      }
    }
  } else if (warp >= 4 * G) {
    // ===================== MMA warp of the group pair (2 pw, 2 pw + 1) =====================
    // Two cursors over the halves of this pair's items.  The P.V cursor follows the softmax groups (O (+)= P_k V_k as
    // soon as P_k is published); the S cursor runs THREE halves ahead (S_{k+3} = Q K_{k+3}^T is issued right after
    // P_k V_k into the TMEM buffer that P_k occupies -- MMAs of one thread execute in issue order, so the overwrite
    // is safe), across item boundaries.  A softmax warp therefore always has its next scores waiting, and the four
    // warps of a group may drift up to two halves apart.  The whole warp runs the control flow; one lane issues.
    ATT2_SETMAXNREG("dec", ATT2_REGS_OTHER);
    const int pw = warp - 4 * G;
    const uint32_t idesc_pv = instr_desc(128, HDP, /*b_mn_major=*/true);
    const uint32_t idesc_s_full = instr_desc(128, 32), idesc_s_last = instr_desc(128, (n_last + 15) & ~15);
    const uint32_t hi_k = (128u >> 4) | (1u << 14);              // K-major tiles: SBO = 128 B
    const uint32_t lo_k = 128u << 16;                            //   128-row tiles: LBO = 128 rows * 16 B
    const uint32_t hi_v = ((128u * 16) >> 4) | (1u << 14);       // V as MN-major: SBO = 2048 B (next 8 columns)
    const uint32_t lo_v = (128u >> 4) << 16;                     //                LBO = 128 B (next 8 kv rows)
    const uint32_t kv16 = (sbase + off_kv) >> 4;
    const uint32_t tile16 = tile_bytes >> 4, stage16 = stage_bytes >> 4;
    const int pv_last_steps = (n_last + 15) >> 4;
    // everything that depends only on the group, once (this loop body runs ~10^4 times per CTA: keep it lean)
    const uint32_t bg0 = BAR(GRP + 16 * (2 * pw)), bg1 = bg0 + 16 * 8;
    const uint32_t tc0 = tmem + (2 * pw) * 128, tc1 = tc0 + 128;
    const uint32_t hs0 = (uint32_t)((2 * pw) / QG) * 2 * tile16, hs1 = (uint32_t)((2 * pw + 1) / QG) * 2 * tile16;
    const uint32_t qt0 = ((sbase + off_q) >> 4) + (2 * pw) * tile16;   // Q tile of group 0 in buffer 0 (16-byte units)
    auto pair_active = [&](int item) {
      int seq, h0, q0;
      decode(item, seq, h0, q0);
      return group_active(2 * pw, h0, q0);
    };
    // ---- S cursor ----
    int s_item = blockIdx.x, s_nl = 0, s_k = 0, s_kq = 0;          // item, its local index, half in the item, half in the stage
    uint32_t s_slot = 0, s_ph = 0, s_bb = 0;                       // ring slot / phase of the half's stage; TMEM buffer
    uint32_t s_q16 = 0, s_st16 = 0;                                // Q tile (group 0) and stage base of the cursor
    int lead = 0;                                                  // halves the S cursor is ahead of the P.V cursor
    int s_stage = 0, pv_stage = 0;                                 // running stage counters of the two cursors
    // S of the cursor's half, in three steps so that the P.V loop can issue it inside its own elected block:
    // s_prepare (waits; false = parked or nothing left), s_emit (MMAs + commits, elected lane only), s_advance.
    auto s_prepare = [&]() -> bool {
      if (s_k == 0) {
        // entering an item.  The cursor never runs past an item this pair has no work in (it parks there until the
        // P.V cursor has done that item's ring / Q bookkeeping): every barrier phase is then observed in order by
        // exactly one of the two cursors, and the ring can never be waited on deeper than it is.
        if (s_item >= p.n_items || !pair_active(s_item)) return false;
        mbar_wait(BAR(Q_FULL + (s_nl & 1)), (uint32_t)(s_nl >> 1) & 1);
        s_q16 = qt0 + (s_nl & 1) * G * tile16;
      }
      if (s_kq == 0) {
        if (s_stage - pv_stage >= NS) return false;                // that ring slot still holds a stage the P.V cursor needs
        mbar_wait(BAR(KV_FULL + s_slot), s_ph);
        s_st16 = kv16 + s_slot * stage16;
      }
      return true;
    };
    auto s_emit = [&]() {
      const bool last = s_k == NH - 1;
      const uint32_t idesc = last ? idesc_s_last : idesc_s_full;
      const uint32_t kb = s_st16 + s_kq * 32, d = s_bb * 32;
#pragma unroll
      for (int kk = 0; kk < KS; ++kk)
        mma_lohi(tc0 + d, (s_q16 + kk * 256) | lo_k, hi_k, (kb + hs0 + kk * 256) | lo_k, hi_k, idesc, (uint32_t)kk);
      mma_commit(bg0 + S_FULL + 8 * s_bb);
#pragma unroll
      for (int kk = 0; kk < KS; ++kk)
        mma_lohi(tc1 + d, (s_q16 + tile16 + kk * 256) | lo_k, hi_k, (kb + hs1 + kk * 256) | lo_k, hi_k, idesc, (uint32_t)kk);
      mma_commit(bg1 + S_FULL + 8 * s_bb);
      if (last) mma_commit(BAR(Q_EMPTY + (s_nl & 1)));             // all S of the item issued: its Q tiles are free after
    };
    auto s_advance = [&]() {
      s_bb = s_bb == 2 ? 0 : s_bb + 1;
      ++s_k; ++lead;
      s_kq = (s_kq + 1) & 3;
      if (s_k == NH) { s_k = 0; s_kq = 0; s_item += gridDim.x; ++s_nl; }
      if (s_kq == 0) { ++s_stage; if (++s_slot == (uint32_t)NS) { s_slot = 0; s_ph ^= 1; } }
    };
    auto issue_next_s = [&]() -> bool {
      if (!s_prepare()) return false;
      tc_fence_after();
      if (elect_one()) s_emit();
      __syncwarp();
      s_advance();
      return true;
    };
    // ---- P.V cursor ----
    uint32_t kslot = 0, kph = 0;
    uint32_t bb = 0, bph = 0;                                      // TMEM buffer of the current half and its phase
    auto release_kv = [&](bool by_commit) {
      if (elect_one()) {
        if (by_commit) mma_commit(BAR(KV_EMPTY + kslot)); else mbar_arrive(BAR(KV_EMPTY + kslot));
      }
      __syncwarp();
      ++pv_stage;
      if (++kslot == (uint32_t)NS) { kslot = 0; kph ^= 1; }
    };
    int n_local = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++n_local) {
      const int b = n_local & 1;
      if (!pair_active(item)) {
        // no work for this pair in the item: keep the ring / Q protocols alive (their barriers expect one arrival
        // per MMA warp, in phase order)
        mbar_wait(BAR(Q_FULL + b), (uint32_t)(n_local >> 1) & 1);
        if (elect_one()) mbar_arrive(BAR(Q_EMPTY + b));
        __syncwarp();
        for (int j = 0; j < NST; ++j) {
          mbar_wait(BAR(KV_FULL + kslot), kph);
          release_kv(false);
          ++s_stage;
          if (++s_slot == (uint32_t)NS) { s_slot = 0; s_ph ^= 1; }   // the S cursor is parked at this item: move it along
        }
        s_item += gridDim.x; ++s_nl;
        continue;
      }
      while (lead < 3 && issue_next_s()) {}
      int kq = 0;
      uint32_t vrow = kv16 + kslot * stage16 + tile16;             // V rows of the current half (head slot 0)
      for (int k = 0; k < NH; ++k) {
        const int n_steps = k == NH - 1 ? pv_last_steps : 2;
        const uint32_t a = bb * 32, pb = 8 * bb;
        const bool do_s = lead <= 3 && s_prepare();                // S_{k+3} rides in the same elected block
        mbar_wait(bg0 + P_FULL + pb, bph);
        mbar_wait(bg1 + P_FULL + pb, bph);
        tc_fence_after();
        if (elect_one()) {
          mma_ts_lohi(tc0 + 96, tc0 + a, (vrow + hs0) | lo_v, hi_v, idesc_pv, (uint32_t)k);
          if (n_steps == 2) mma_ts_lohi(tc0 + 96, tc0 + a + 8, (vrow + hs0 + 16) | lo_v, hi_v, idesc_pv, 1u);
          mma_commit(bg0 + PV_DONE + pb);
          mma_ts_lohi(tc1 + 96, tc1 + a, (vrow + hs1) | lo_v, hi_v, idesc_pv, (uint32_t)k);
          if (n_steps == 2) mma_ts_lohi(tc1 + 96, tc1 + a + 8, (vrow + hs1 + 16) | lo_v, hi_v, idesc_pv, 1u);
          mma_commit(bg1 + PV_DONE + pb);
          if (do_s) s_emit();
        }
        __syncwarp();
        --lead;
        if (do_s) s_advance();
        while (lead < 3 && issue_next_s()) {}
        if (bb == 2) { bb = 0; bph ^= 1; } else ++bb;
        vrow += 32;
        if (++kq == 4 || k == NH - 1) {                            // this pair is done with the K/V stage
          release_kv(true);
          kq = 0;
          vrow = kv16 + kslot * stage16 + tile16;
        }
      }
    }
  } else {
    // ===================== softmax groups =====================
    ATT2_SETMAXNREG("inc", ATT2_REGS_SOFTMAX);
    const int g = warp >> 2;
    const int quarter = warp & 3;                                  // TMEM lane quarter this warp may access
    const int m = quarter * 32 + lane;
    const uint32_t tcol = tmem + ((uint32_t)(quarter * 32) << 16) + g * 128;
    const uint32_t bgrp = BAR(GRP + 16 * g);
    const int OC = HDP / 8;
    uint32_t kk = 0, ph = 0;                                       // running half counter of this group, phase of its buffer
    int bb = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      int seq, h0, q0;
      decode(item, seq, h0, q0);
      if (!group_active(g & ~1, h0, q0)) continue;                 // the pair has no work in this item
      const bool store = group_active(g, h0, q0);
      const int head = h0 + g / QG, qt = q0 + g % QG;
      float m_ref = 0.f, l0 = 0.f, l1 = 0.f;
      const int n_full = n_last == 32 ? NH : NH - 1;
      int lb = 0; uint32_t lph = 0;                                // buffer / phase of the item's last half
#if ATT2_PREFETCH
      // Two score registers sets: the read-back of half k + 1 is issued (after its S_FULL) before the exp2 pass of
      // half k, so TMEM-read time hides behind MUFU time inside the warp.  Loop unrolled by two to keep both sets in
      // registers.  (bb, ph) describe half k; (nb, nph) half k + 1.
      uint32_t sA[32], sB[32];
      mbar_wait(bgrp + S_FULL + 8 * bb, ph);
      tc_fence_after();
      tmem_ld32(tcol + bb * 32, sA);
#define ATT2_STEP(CUR, NXT)                                                                                         \
      {                                                                                                             \
        tc_wait_ld();                                                                                               \
        const int nb = bb == 2 ? 0 : bb + 1;                                                                        \
        const uint32_t nph = bb == 2 ? ph ^ 1 : ph;                                                                 \
        if (k + 1 < NH) {                                                                                           \
          mbar_wait(bgrp + S_FULL + 8 * nb, nph);                                                                   \
          tc_fence_after();                                                                                         \
          tmem_ld32(tcol + nb * 32, NXT);                                                                           \
        }                                                                                                           \
        if (k < n_full) attn2_half<true>(CUR, tcol, bgrp, bb, ph, kk, 32, k == 0, HDP, m_ref, l0, l1);              \
        else attn2_half<false>(CUR, tcol, bgrp, bb, ph, kk, n_last, k == 0, HDP, m_ref, l0, l1);                    \
        tc_fence_before();                                                                                          \
        mbar_arrive(bgrp + P_FULL + 8 * bb);                                                                        \
        lb = bb; lph = ph;                                                                                          \
        bb = nb; ph = nph;                                                                                          \
        ++k; ++kk;                                                                                                  \
      }
      for (int k = 0; k < NH;) {
        ATT2_STEP(sA, sB)
        if (k < NH) ATT2_STEP(sB, sA)
      }
#undef ATT2_STEP
#elif ATT2_QUARTER
      uint32_t sA[16], sB[16];
      mbar_wait(bgrp + S_FULL + 8 * bb, ph);
      tc_fence_after();
      tmem_ld16(tcol + bb * 32, sA);                               // first quarter of the item's first half
      for (int k = 0; k < NH; ++k, ++kk) {
        const int nk = k < n_full ? 32 : n_last;                   // keys in this half
        tc_wait_ld();                                              // sA: scores 0..15 of half k
        tmem_ld16(tcol + bb * 32 + 16, sB);                        // scores 16..31 in flight during the first exp2 pass
        if (nk < 16) {
#pragma unroll
          for (int i = 0; i < 16; ++i) if (i >= nk) sA[i] = 0xff800000u;
        }
        const float mxa = attn2_max16(sA);
        if (k == 0) {
          m_ref = mxa;
        } else if (__any_sync(0xffffffffu, mxa > m_ref + ATT2_TH)) {
          attn2_raise(tcol, bgrp, bb, ph, true, HDP, mxa > m_ref + ATT2_TH ? mxa : m_ref, m_ref, l0, l1);
        }
        if (kk >= 3) mbar_wait(bgrp + PV_DONE + 8 * bb, ph ^ 1);   // phase bookkeeping only (see attn2_half)
        uint32_t pk[8];
        attn2_exp16<true>(sA, m_ref, pk, l0, l1);
        tmem_st8(tcol + bb * 32, pk);
        tc_wait_ld();                                              // sB
        if (nk < 32) {
#pragma unroll
          for (int i = 0; i < 16; ++i) if (16 + i >= nk) sB[i] = 0xff800000u;
        }
        const float mxb = attn2_max16(sB);
        if (__any_sync(0xffffffffu, mxb > m_ref + ATT2_TH)) {
          // the second quarter raises the reference: the first quarter's probabilities were written against the old one
          attn2_raise(tcol, bgrp, bb, ph, k > 0, HDP, mxb > m_ref + ATT2_TH ? mxb : m_ref, m_ref, l0, l1);
          tc_wait_st();
          attn2_exp16<false>(sA, m_ref, pk, l0, l1);
          tmem_st8(tcol + bb * 32, pk);
        }
        const int nb = bb == 2 ? 0 : bb + 1;
        const uint32_t nph = bb == 2 ? ph ^ 1 : ph;
        if (k + 1 < NH) {                                          // sA is dead: first quarter of the next half
          mbar_wait(bgrp + S_FULL + 8 * nb, nph);
          tc_fence_after();
          tmem_ld16(tcol + nb * 32, sA);
        }
        if (nk > 16) {
          attn2_exp16<true>(sB, m_ref, pk, l0, l1);
          tmem_st8(tcol + bb * 32 + 8, pk);
        }
        tc_wait_st();
        tc_fence_before();
        mbar_arrive(bgrp + P_FULL + 8 * bb);
        lb = bb; lph = ph;
        bb = nb; ph = nph;
      }
#else
      for (int k = 0; k < NH; ++k, ++kk) {
        mbar_wait(bgrp + S_FULL + 8 * bb, ph);
        tc_fence_after();
        uint32_t s[32];
        tmem_ld32(tcol + bb * 32, s);
        tc_wait_ld();
        if (k < n_full) attn2_half<true>(s, tcol, bgrp, bb, ph, kk, 32, k == 0, HDP, m_ref, l0, l1);
        else attn2_half<false>(s, tcol, bgrp, bb, ph, kk, n_last, k == 0, HDP, m_ref, l0, l1);
        tc_fence_before();
        mbar_arrive(bgrp + P_FULL + 8 * bb);
        lb = bb; lph = ph;
        if (++bb == 3) { bb = 0; ph ^= 1; }
      }
#endif
      // ---- all keys done: O / l -> this head's slice of the o image ----
      mbar_wait(bgrp + PV_DONE + 8 * lb, lph);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld16(tcol + 96, r);
      if (HDP > 16) tmem_ld16(tcol + 96 + 16, r + 16);
      tc_wait_ld();
      if (store) {
        const float inv = 1.f / (l0 + l1);
        __nv_bfloat16* ob = p.o + (((size_t)seq * NTL + qt) * (p.heads * OC) + (size_t)head * OC) * 1024 + (size_t)m * 8;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < OC) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              w[e] = pack_bf16(__uint_as_float(r[c * 8 + 2 * e]) * inv, __uint_as_float(r[c * 8 + 2 * e + 1]) * inv);
            *reinterpret_cast<uint4*>(ob + (size_t)c * 1024) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4 * G) tmem_dealloc(tmem, 512);
}

}  // namespace tfl
