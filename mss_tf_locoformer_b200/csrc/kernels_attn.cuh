// K5  attn_tc_kernel: softmax(q k^T) v per (sequence, head) on tcgen05 / TMEM
// (models/mss_tflocoformer.py:523-531).  q arrives RoPE-rotated and pre-scaled by
// log2(e)/sqrt(hd), so the softmax is exp2(s - max).
//
// Operands live in HBM as 128-row "tile images" in the chunk-major shared-memory layout of
// tc_common.cuh, so every tile is one contiguous block fetched by a single bulk async copy:
//   qkv image : [which q|k|v][seq][head][tile jt][HDP/8 chunks][128 rows][8]  bf16
//   o   image : [seq][tile jt][heads * HDP / 8 chunks][128 rows][8]            bf16 (A operand of the
//               head-merge projection, heads side by side)
// HDP = head_dim rounded up to 16 (zero columns); rows >= L of the last tile are zero.
//
// One persistent CTA per SM, work item = (seq, head, pair of 128-row query tiles):
//   warp 0      loader: Q tiles, K/V ring (one stage = K tile + V tile)
//   warp 1      MMA thread: S[g] = Q[g] K^T (TMEM, 128 cols), O_j[g] = P[g] V (TMEM, HDP cols, fresh per tile)
//   warps 2-9   two softmax groups (one per query tile): TMEM -> exp2 -> bf16 P tile in smem (A operand),
//               running max / sum and the output accumulator (rescaled per tile) stay in registers.
// With head_dim 32 the kernel is bound by the 16/clk/SM exp2 unit, not by the tensor pipe.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace tfl {

struct AttnTcParams {
  const __nv_bfloat16* qkv; __nv_bfloat16* o;
  int nseq, heads, L, NTL, HDP, NP;   // NTL = 128-row tiles of the images; NP = query-tile pairs per (seq, head)
  int NQT;                            // query tiles handled here (the last rows may go to attn_tail_rows_kernel)
  int NU;                             // 64-key units per sequence (the last one may be partial: masked)
  int n_items;
};

constexpr int ATT_STAGES = 4;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(352, 1) attn_tc_kernel(AttnTcParams p) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HDP = p.HDP, NTL = p.NTL;
  const int NU = p.NU;                                            // 64-key units per sequence
  const int NST = (NU + 1) / 2;                                   // K/V ring stages (128-row tiles) per item
  const uint32_t tile_bytes = (uint32_t)HDP * 128 * 2;          // one Q / K / V tile (128 rows)
  const uint32_t p_bytes = 128u * 64 * 2;                        // one P half-tile (128 queries x 64 keys)
  const uint32_t off_q = 0, off_kv = 4 * tile_bytes, off_p = off_kv + ATT_STAGES * 2 * tile_bytes;
  const uint32_t off_bar = off_p + 4 * p_bytes;
  const uint32_t sbase = smem_u32(smem);
  auto BAR = [&](int i) { return sbase + off_bar + 8u * i; };
  // 0..3 kv_full, 4..7 kv_empty, 8..11 q_full[g][item parity], 12..15 q_empty, then [g*2 + buf] families:
  // 16 s_full, 20 s_empty, 24 p_full, 28 p_empty, 32 o_full, 36 o_empty; slot 44: TMEM base
  const int KV_FULL = 0, KV_EMPTY = 4, Q_FULL = 8, Q_EMPTY = 12, S_FULL = 16, S_EMPTY = 20, P_FULL = 24, P_EMPTY = 28,
            O_FULL = 32, O_EMPTY = 36;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + off_bar + 8 * 44);
  if (threadIdx.x == 0) {
    for (int i = 0; i < ATT_STAGES; ++i) { mbar_init(BAR(KV_FULL + i), 1); mbar_init(BAR(KV_EMPTY + i), 2); }
    for (int i = 0; i < 4; ++i) {
      mbar_init(BAR(Q_FULL + i), 1); mbar_init(BAR(Q_EMPTY + i), 1);
      mbar_init(BAR(S_FULL + i), 1); mbar_init(BAR(S_EMPTY + i), 128);
      mbar_init(BAR(P_FULL + i), 128); mbar_init(BAR(P_EMPTY + i), 1);
      mbar_init(BAR(O_FULL + i), 1); mbar_init(BAR(O_EMPTY + i), 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns: S[g][buf] at (g*2+buf)*64 (4 x 64), O_u[g][buf] at 256 + (g*2+buf)*HDP
  const uint32_t o_col0 = 256;
  const size_t which_stride = (size_t)p.nseq * p.heads * NTL * HDP * 128;   // elements between q, k, v planes

  if (warp == 0) {
    // ===================== loader (whole warp runs the loop, one elected lane issues) =====================
    {
      uint32_t kslot = 0, kph = 0, qph = 0;    // qph: bit (g*2 + b) = phase of Q buffer [g][b]
      int n_local = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++n_local) {
        const int pair = item % p.NP, sh = item / p.NP;            // sh = seq * heads + head
        const __nv_bfloat16* qb = p.qkv + (size_t)sh * NTL * HDP * 128;
        const int nq = min(2, p.NQT - 2 * pair);
        const int b = n_local & 1;
        for (int g = 0; g < nq; ++g) {
          const int qi = g * 2 + b;
          mbar_wait(BAR(Q_EMPTY + qi), ((qph >> qi) & 1) ^ 1);
          qph ^= 1u << qi;
          if (elect_one()) {
            mbar_arrive_expect_tx(BAR(Q_FULL + qi), tile_bytes);
            bulk_g2s(sbase + off_q + qi * tile_bytes, qb + (size_t)(2 * pair + g) * HDP * 128, tile_bytes, BAR(Q_FULL + qi));
          }
          __syncwarp();
        }
        for (int j = 0; j < NST; ++j) {
          mbar_wait(BAR(KV_EMPTY + kslot), kph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(BAR(KV_FULL + kslot), 2 * tile_bytes);
            const uint32_t dst = sbase + off_kv + kslot * 2 * tile_bytes;
            bulk_g2s(dst, qb + which_stride + (size_t)j * HDP * 128, tile_bytes, BAR(KV_FULL + kslot));
            bulk_g2s(dst + tile_bytes, qb + 2 * which_stride + (size_t)j * HDP * 128, tile_bytes, BAR(KV_FULL + kslot));
          }
          __syncwarp();
          if (++kslot == ATT_STAGES) { kslot = 0; kph ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ===================== MMA issuers: one warp per query-tile group =====================
    // Work is cut into units of 64 keys (half a K/V tile).  S for unit u+1 (next half tile, or unit 0 of the next
    // item) is issued BEFORE P.V of unit u and S / P / O are double-buffered, so a softmax group never waits on the
    // tensor pipe in steady state.  The whole warp runs the control flow; one elected lane issues.
    // Ordering that makes extra barriers unnecessary: the softmax group folds O(u-1) before it publishes P(u+1),
    // so "P(u) full" implies O[buf(u)] has been read out, and "O(u-2) full" (awaited by fold) implies P[buf(u)] is free.
    {
      const int g = warp == 1 ? 0 : 1;
      const uint32_t idesc_s = instr_desc(128, 64), idesc_pv = instr_desc(128, HDP, /*b_mn_major=*/true);
      const uint32_t hi_k = (128u >> 4) | (1u << 14);              // K-major tiles: SBO = 128 B
      const uint32_t lo_k = 128u << 16;                            //   128-row tiles: LBO = 128 rows * 16 B
      const uint32_t hi_v = ((128u * 16) >> 4) | (1u << 14);       // V as MN-major: SBO = 2048 B (next 8 columns)
      const uint32_t lo_v = (128u >> 4) << 16;                     //                LBO = 128 B (next 8 kv rows)
      const uint32_t q16 = (sbase + off_q) >> 4, kv16 = (sbase + off_kv) >> 4, p16 = (sbase + off_p) >> 4;
      const uint32_t tile16 = tile_bytes >> 4, pt16 = p_bytes >> 4;
      uint32_t kslot = 0, kph = 0;
      uint32_t qph = 0, sph = 0, pph = 0;   // phase bits indexed by buffer
      uint32_t us = 0, up = 0;              // running unit counters (S-issue side, P.V side)
      auto issue_s = [&](int b, uint32_t slot, int half, bool last_of_item) {
        const int buf = (int)(us & 1), si = g * 2 + buf;
        ++us;
        mbar_wait(BAR(S_EMPTY + si), ((sph >> buf) & 1) ^ 1);
        sph ^= 1u << buf;
        tc_fence_after();
        const uint32_t qa = q16 + (g * 2 + b) * tile16, kb = kv16 + slot * 2 * tile16 + half * 64;
        if (elect_one()) {
          for (int kk = 0; kk < HDP / 16; ++kk)
            mma_lohi(tmem + si * 64, (qa + kk * 2 * 128) | lo_k, hi_k, (kb + kk * 2 * 128) | lo_k, hi_k, idesc_s, (uint32_t)kk);
          mma_commit(BAR(S_FULL + si));
          if (last_of_item) mma_commit(BAR(Q_EMPTY + g * 2 + b));
        }
        __syncwarp();
      };
      auto wait_q = [&](int b) {
        mbar_wait(BAR(Q_FULL + g * 2 + b), (qph >> b) & 1);
        qph ^= 1u << b;
      };
      auto release_kv = [&]() {
        if (elect_one()) mma_commit(BAR(KV_EMPTY + kslot));
        __syncwarp();
        if (++kslot == ATT_STAGES) { kslot = 0; kph ^= 1; }
      };
      int n_local = 0;
      if ((int)blockIdx.x < p.n_items && g < min(2, p.NQT - 2 * ((int)blockIdx.x % p.NP))) {   // S of the very first unit
        wait_q(0);
        mbar_wait(BAR(KV_FULL + kslot), kph);
        tc_fence_after();
        issue_s(0, kslot, 0, NU == 1);
      }
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++n_local) {
        const bool active = g < min(2, p.NQT - 2 * (item % p.NP));
        const int b = n_local & 1;
        const int next = item + gridDim.x;
        const bool next_active = next < p.n_items && g < min(2, p.NQT - 2 * (next % p.NP));
        if (!active) {
          // this group has no query tile in the item: keep the K/V ring protocol alive (the stage barriers expect
          // one release from each MMA warp) and pre-issue S for the next item after the last stage
          for (int j = 0; j < NST; ++j) {
            mbar_wait(BAR(KV_FULL + kslot), kph);
            if (j + 1 == NST && next_active) {
              uint32_t nslot = kslot + 1, nph = kph;
              if (nslot == ATT_STAGES) { nslot = 0; nph ^= 1; }
              wait_q(b ^ 1);
              mbar_wait(BAR(KV_FULL + nslot), nph);
              tc_fence_after();
              issue_s(b ^ 1, nslot, 0, NU == 1);
            }
            release_kv();
          }
          continue;
        }
        for (int u = 0; u < NU; ++u) {
          const int half = u & 1;
          uint32_t nslot = kslot + 1, nph = kph;
          if (nslot == ATT_STAGES) { nslot = 0; nph ^= 1; }
          // ---- S for the next unit ----
          if (u + 1 < NU) {
            if (half == 1) { mbar_wait(BAR(KV_FULL + nslot), nph); tc_fence_after(); }
            issue_s(b, half == 1 ? nslot : kslot, half ^ 1, u + 2 == NU);
          } else if (next_active) {
            wait_q(b ^ 1);
            mbar_wait(BAR(KV_FULL + nslot), nph);
            tc_fence_after();
            issue_s(b ^ 1, nslot, 0, NU == 1);
          }
          // ---- O_u = P_u . V[64 keys of this unit] ----
          const int buf = (int)(up & 1), bi = g * 2 + buf;
          ++up;
          mbar_wait(BAR(P_FULL + bi), (pph >> buf) & 1);
          pph ^= 1u << buf;
          tc_fence_after();
          const uint32_t pa = p16 + bi * pt16, vb = kv16 + kslot * 2 * tile16 + tile16 + half * 64;
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              mma_lohi(tmem + o_col0 + bi * HDP, (pa + kk * 2 * 128) | lo_k, hi_k, (vb + kk * 16) | lo_v, hi_v, idesc_pv,
                       (uint32_t)kk);
            mma_commit(BAR(O_FULL + bi));
          }
          __syncwarp();
          if (half == 1 || u + 1 == NU) release_kv();   // both halves of this K/V stage consumed by this group
        }
      }
    }
  } else {
    // ===================== softmax groups =====================
    const int g = (warp - 2) >> 2;              // warps 2-5: group 0, warps 6-9: group 1
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    uint32_t sph = 0, oph = 0;                  // phase bits per buffer (bit = buf)
    uint32_t uc = 0;                            // running unit counter of this group
    const int OC = HDP / 8;                     // 16-byte chunks per output row and head
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int pair = item % p.NP, sh = item / p.NP;
      const int nq = min(2, p.NQT - 2 * pair);
      if (g >= nq) continue;
      const int qt = 2 * pair + g;
      float m_run = -INFINITY, l_run = 0.f;
      float o[32];
#pragma unroll
      for (int d = 0; d < 32; ++d) o[d] = 0.f;
      auto fold = [&](int ob) {               // o += O of the previous unit
        const int bi = g * 2 + ob;
        mbar_wait(BAR(O_FULL + bi), (oph >> ob) & 1);
        oph ^= 1u << ob;
        tc_fence_after();
        uint32_t r[32];
        tmem_ld16(lane_addr + o_col0 + bi * HDP, r);
        if (HDP > 16) tmem_ld16(lane_addr + o_col0 + bi * HDP + 16, r + 16);
        tc_wait_ld();
#pragma unroll
        for (int e = 0; e < 32; ++e) if (e < HDP) o[e] += __uint_as_float(r[e]);
      };
      for (int u = 0; u < NU; ++u, ++uc) {
        const int buf = (int)(uc & 1), bi = g * 2 + buf;
        mbar_wait(BAR(S_FULL + bi), (sph >> buf) & 1);
        sph ^= 1u << buf;
        tc_fence_after();
        uint32_t s[64];
        tmem_ld32(lane_addr + bi * 64, s);
        tmem_ld32(lane_addr + bi * 64 + 32, s + 32);
        tc_wait_ld();
        tc_fence_before();
        mbar_arrive(BAR(S_EMPTY + bi));
        const int valid = p.L - u * 64;
        if (valid < 64) {
#pragma unroll
          for (int i = 0; i < 64; ++i) if (i >= valid) s[i] = 0xff800000u;  // -inf: key rows beyond the sequence
        }
        float mx0 = m_run, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(s[i])); mx1 = fmaxf(mx1, __uint_as_float(s[i + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(s[i + 2])); mx3 = fmaxf(mx3, __uint_as_float(s[i + 3]));
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        const float alpha = fast_exp2(m_run - mx);
        m_run = mx;
        float rs0 = 0.f, rs1 = 0.f;
        // P[buf] is free: fold() of the previous iteration waited for P.V(u-2), the last reader of this buffer
        uint8_t* pt = smem + off_p + (size_t)bi * p_bytes;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float p0 = fast_exp2(__uint_as_float(s[c * 8 + 2 * e]) - mx);
            const float p1 = fast_exp2(__uint_as_float(s[c * 8 + 2 * e + 1]) - mx);
            rs0 += p0; rs1 += p1;
            w[e] = pack_bf16(p0, p1);
          }
          *reinterpret_cast<uint4*>(pt + ((size_t)c * 128 + m) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(BAR(P_FULL + bi));
        if (u > 0) fold(buf ^ 1);
        l_run = l_run * alpha + (rs0 + rs1);
#pragma unroll
        for (int d = 0; d < 32; ++d) o[d] *= alpha;
      }
      fold((int)((uc - 1) & 1));
      const int s_idx = sh / p.heads, h = sh - s_idx * p.heads;
      // ---- normalise and store this head's slice of the o image ----
      const float inv = 1.f / l_run;
      __nv_bfloat16* ob = p.o + (((size_t)s_idx * NTL + qt) * (p.heads * OC) + (size_t)h * OC) * 1024 + (size_t)m * 8;
      for (int c = 0; c < OC; ++c) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) w[e] = pack_bf16(o[(c * 8 + 2 * e) & 31] * inv, o[(c * 8 + 2 * e + 1) & 31] * inv);
        *reinterpret_cast<uint4*>(ob + (size_t)c * 1024) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// Rows of a sequence that do not fill a 128-row query tile (L mod 128 <= 8: e.g. 1025 = 8*128 + 1, 259 = 2*128 + 3)
// would each cost a whole tensor-core tile; they run here instead.  One warp per (sequence, head, group of RG tail
// rows): lanes over keys -- every K and V row is fetched once and used for all RG rows --, shuffle softmax, shuffle
// reduction of the output.  Same bf16 images, same exp2 domain as the tensor-core kernel.
template <int RG>
__global__ void __launch_bounds__(256) attn_tail_rows_kernel(AttnTcParams p, int row0) {
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const int n_tail = p.L - row0, n_grp = (n_tail + RG - 1) / RG;
  const long long total = (long long)p.nseq * p.heads * n_grp;
  const int HDP = p.HDP, NTL = p.NTL, OC = HDP / 8;
  const size_t plane = (size_t)p.nseq * p.heads * NTL * HDP * 128;
  extern __shared__ float tail_sm[];                     // per warp: scores of RG rows x Lpad keys
  const int Lpad = (p.L + 31) & ~31;
  float* prob = tail_sm + (size_t)(threadIdx.x >> 5) * RG * Lpad;
  for (long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < total;
       w += (long long)gridDim.x * (blockDim.x >> 5)) {
    const int grp = (int)(w % n_grp), sh = (int)(w / n_grp);
    const int r0 = row0 + grp * RG;
    const __nv_bfloat16* base = p.qkv + (size_t)sh * NTL * HDP * 128;
    float qf[RG][32];
#pragma unroll
    for (int r = 0; r < RG; ++r) {
      const int row = min(r0 + r, p.L - 1);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 raw = make_uint4(0u, 0u, 0u, 0u);
        if (c < OC) raw = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(row >> 7) * HDP * 128 + ((size_t)c * 128 + (row & 127)) * 8));
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h2[e]); qf[r][c * 8 + 2 * e] = f.x; qf[r][c * 8 + 2 * e + 1] = f.y; }
      }
    }
    float mx[RG];
#pragma unroll
    for (int r = 0; r < RG; ++r) mx[r] = -INFINITY;
    for (int j = lane; j < Lpad; j += 32) {
      float sc[RG];
#pragma unroll
      for (int r = 0; r < RG; ++r) sc[r] = j < p.L ? 0.f : -INFINITY;
      if (j < p.L) {
        const __nv_bfloat16* krow = base + plane + (size_t)(j >> 7) * HDP * 128 + (size_t)(j & 127) * 8;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < OC) {
            const uint4 kr = __ldg(reinterpret_cast<const uint4*>(krow + (size_t)c * 1024));
            const __nv_bfloat162* k2 = reinterpret_cast<const __nv_bfloat162*>(&kr);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 kf = __bfloat1622float2(k2[e]);
#pragma unroll
              for (int r = 0; r < RG; ++r) sc[r] = fmaf(qf[r][c * 8 + 2 * e], kf.x, fmaf(qf[r][c * 8 + 2 * e + 1], kf.y, sc[r]));
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < RG; ++r) { prob[r * Lpad + j] = sc[r]; mx[r] = fmaxf(mx[r], sc[r]); }
    }
#pragma unroll
    for (int r = 0; r < RG; ++r)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], o));
    // exp2, row sums and P.V in one pass over the keys
    float sum[RG], acc[RG][32];
#pragma unroll
    for (int r = 0; r < RG; ++r) {
      sum[r] = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) acc[r][d] = 0.f;
    }
    for (int j = lane; j < p.L; j += 32) {
      float pj[RG];
#pragma unroll
      for (int r = 0; r < RG; ++r) { pj[r] = fast_exp2(prob[r * Lpad + j] - mx[r]); sum[r] += pj[r]; }
      const __nv_bfloat16* vrow = base + 2 * plane + (size_t)(j >> 7) * HDP * 128 + (size_t)(j & 127) * 8;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < OC) {
          const uint4 vr = __ldg(reinterpret_cast<const uint4*>(vrow + (size_t)c * 1024));
          const __nv_bfloat162* v2 = reinterpret_cast<const __nv_bfloat162*>(&vr);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 vf = __bfloat1622float2(v2[e]);
#pragma unroll
            for (int r = 0; r < RG; ++r) {
              acc[r][c * 8 + 2 * e] = fmaf(pj[r], vf.x, acc[r][c * 8 + 2 * e]);
              acc[r][c * 8 + 2 * e + 1] = fmaf(pj[r], vf.y, acc[r][c * 8 + 2 * e + 1]);
            }
          }
        }
      }
    }
    const int s_idx = sh / p.heads, h = sh - s_idx * p.heads;
#pragma unroll
    for (int r = 0; r < RG; ++r) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], o);
#pragma unroll
        for (int d = 0; d < 32; ++d) acc[r][d] += __shfl_xor_sync(0xffffffffu, acc[r][d], o);
      }
      float mine = 0.f;
#pragma unroll
      for (int d = 0; d < 32; ++d) if (d == lane) mine = acc[r][d];
      const int row = r0 + r;
      if (lane < HDP && row < p.L)
        p.o[(((size_t)s_idx * NTL + (row >> 7)) * (p.heads * OC) + (size_t)h * OC + (lane >> 3)) * 1024 + (size_t)(row & 127) * 8 + (lane & 7)] =
            __float2bfloat16_rn(mine / sum[r]);
    }
    __syncwarp();
  }
}

// Tail rows on the warp-level tensor-core path (mma.sync m16n8k16, bf16 x bf16 -> fp32): the CUDA-core kernel above
// is bound by its instruction count on the time axis (3 tail rows x 4 heads x 8200 sequences: ~10 k warp instructions
// per (sequence, head) for the scores, the bf16 -> fp32 conversions and the P.V FMAs; ncu r02: 0.47 ms per call against
// 0.17 ms of K / V traffic).  Here one warp owns a (sequence, head): the <= 8 tail rows are rows 0..7 of an M = 16 tile,
// keys are walked 16 at a time with an online softmax (flash style), ~60 instructions per 16 keys for ALL rows.
//   S  = Q K^T : A = Q rows (fragments straight from the q image: 4-byte loads), B = K rows (4-byte loads, one 128-byte
//                line per warp load)
//   O += P V   : A = P (the S accumulator fragments re-packed as bf16: C layout of two n-tiles == A layout of one k-step),
//                B = V via ldmatrix.trans from a 16-key block staged in shared memory (V is stored dims-contiguous).
// Same images, same exp2 domain (q carries log2(e) / sqrt(hd)), P rounded to bf16 as in attn_tc2_kernel.
template <int KS>   // head_dim padded to 16 * KS
__global__ void __launch_bounds__(256) attn_tail_mma_kernel(AttnTcParams p, int row0) {
  constexpr int HDP = 16 * KS, NT = 2 * KS;                  // NT = 8-wide n-tiles of the head dimension
  constexpr int VPITCH = HDP * 2 + 16;                       // bytes per staged V row (padded: conflict-free ldmatrix)
  __shared__ __align__(16) uint8_t vstage[8][16 * VPITCH];
  griddep_wait();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int n_tail = p.L - row0, NTL = p.NTL, OC = HDP / 8;
  const long long total = (long long)p.nseq * p.heads;
  const size_t plane = (size_t)p.nseq * p.heads * NTL * HDP * 128;
  uint8_t* vs = vstage[wib];
  const uint32_t vs_addr = (uint32_t)__cvta_generic_to_shared(vs);
  for (long long sh = (long long)blockIdx.x * 8 + wib; sh < total; sh += (long long)gridDim.x * 8) {
    const __nv_bfloat16* qimg = p.qkv + (size_t)sh * NTL * HDP * 128;
    const __nv_bfloat16* kimg = qimg + plane;
    const __nv_bfloat16* vimg = qimg + 2 * plane;
    // ---- Q fragments (rows g < n_tail; rows 8..15 of the tile stay zero) ----
    uint32_t qa[KS][2];
    {
      const int row = row0 + g;
      const bool valid = g < n_tail;
      const __nv_bfloat16* qt = qimg + (size_t)(min(row, p.L - 1) >> 7) * HDP * 128 + (size_t)(min(row, p.L - 1) & 127) * 8 + 2 * t;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        qa[ks][0] = valid ? __ldg(reinterpret_cast<const uint32_t*>(qt + (size_t)(2 * ks) * 1024)) : 0u;
        qa[ks][1] = valid ? __ldg(reinterpret_cast<const uint32_t*>(qt + (size_t)(2 * ks + 1) * 1024)) : 0u;
      }
    }
    float o[NT][2];
#pragma unroll
    for (int j = 0; j < NT; ++j) o[j][0] = o[j][1] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    // K fragments and V pieces of the NEXT 16-key block are requested before the current block is processed (one
    // global round trip per block would otherwise be exposed: 65 of them for a 1025-bin sequence)
    constexpr int VP = (2 * KS * 16 + 31) / 32;               // 16-byte V pieces per lane and block
    uint32_t kf[2][KS][2], kf_n[2][KS][2];
    uint4 vp[VP], vp_n[VP];
    auto fetch = [&](int k0, uint32_t (&kq)[2][KS][2], uint4 (&vq)[VP]) {
      const size_t tile_off = (size_t)(k0 >> 7) * HDP * 128;
      const int kr = k0 & 127;
#pragma unroll
      for (int i = 0; i < VP; ++i) {
        const int piece = lane + 32 * i;
        const int c = piece >> 4, key = piece & 15;
        vq[i] = piece < OC * 16 ? __ldg(reinterpret_cast<const uint4*>(vimg + tile_off + ((size_t)c * 128 + kr + key) * 8))
                                : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const __nv_bfloat16* kt = kimg + tile_off + (size_t)(kr + 8 * n + g) * 8 + 2 * t;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          kq[n][ks][0] = __ldg(reinterpret_cast<const uint32_t*>(kt + (size_t)(2 * ks) * 1024));
          kq[n][ks][1] = __ldg(reinterpret_cast<const uint32_t*>(kt + (size_t)(2 * ks + 1) * 1024));
        }
      }
    };
    fetch(0, kf, vp);
    for (int k0 = 0; k0 < p.L; k0 += 16) {
      if (k0 + 16 < p.L) fetch(k0 + 16, kf_n, vp_n);
      // stage the 16 V rows of this block: OC chunks x 16 keys x 16 B
#pragma unroll
      for (int i = 0; i < VP; ++i) {
        const int piece = lane + 32 * i;
        if (piece < OC * 16) *reinterpret_cast<uint4*>(vs + (piece & 15) * VPITCH + (piece >> 4) * 16) = vp[i];
      }
      // ---- S = Q K^T for keys k0 .. k0 + 15 (two n-tiles of 8 keys) ----
      float sc[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(sc[n][0]), "+f"(sc[n][1]), "+f"(sc[n][2]), "+f"(sc[n][3])
                       : "r"(qa[ks][0]), "r"(0u), "r"(qa[ks][1]), "r"(0u), "r"(kf[n][ks][0]), "r"(kf[n][ks][1]));
        }
      }
      // ---- online softmax on row g (this lane holds keys k0 + 8n + 2t, +1) ----
      float s4[4] = {sc[0][0], sc[0][1], sc[1][0], sc[1][1]};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int key = k0 + 8 * (i >> 1) + 2 * t + (i & 1);
        if (key >= p.L) s4[i] = -INFINITY;
      }
      float mx = fmaxf(fmaxf(s4[0], s4[1]), fmaxf(s4[2], s4[3]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run, mx);                  // finite: every block holds at least one key < L
      const float alpha = fast_exp2(m_run - m_new);
      float pr[4], psum = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) { pr[i] = fast_exp2(s4[i] - m_new); psum += pr[i]; }
      l_run = l_run * alpha + psum;
      m_run = m_new;
#pragma unroll
      for (int j = 0; j < NT; ++j) { o[j][0] *= alpha; o[j][1] *= alpha; }
      const uint32_t pa0 = tc::pack_bf16(pr[0], pr[1]), pa2 = tc::pack_bf16(pr[2], pr[3]);   // A fragment rows g: k = 2t.., 2t + 8..
      // ---- O += P V: B fragments through ldmatrix.trans (two n-tiles of dims per x4) ----
      __syncwarp();
#pragma unroll
      for (int jp = 0; jp < NT / 2; ++jp) {
        const int mtx = lane >> 3, r = lane & 7;             // matrix m: keys (m & 1) * 8 + r, dims (2 jp + (m >> 1)) * 8
        const uint32_t addr = vs_addr + ((mtx & 1) * 8 + r) * VPITCH + (2 * jp + (mtx >> 1)) * 16;
        uint32_t b[4];
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "r"(addr));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float d2 = 0.f, d3 = 0.f;                          // rows 8..15 of the tile: unused
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(o[2 * jp + h][0]), "+f"(o[2 * jp + h][1]), "+f"(d2), "+f"(d3)
                       : "r"(pa0), "r"(0u), "r"(pa2), "r"(0u), "r"(b[2 * h]), "r"(b[2 * h + 1]));
        }
      }
      __syncwarp();                                          // the staging rows are rewritten by the next block
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) { kf[n][ks][0] = kf_n[n][ks][0]; kf[n][ks][1] = kf_n[n][ks][1]; }
#pragma unroll
      for (int i = 0; i < VP; ++i) vp[i] = vp_n[i];
    }
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
    l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
    if (g < n_tail) {
      const float inv = 1.f / l_run;
      const int row = row0 + g;
      const int s_idx = (int)(sh / p.heads), h = (int)(sh - (long long)s_idx * p.heads);
      __nv_bfloat16* ob = p.o + (((size_t)s_idx * NTL + (row >> 7)) * (p.heads * OC) + (size_t)h * OC) * 1024 +
                          (size_t)(row & 127) * 8 + 2 * t;
#pragma unroll
      for (int j = 0; j < NT; ++j)
        *reinterpret_cast<uint32_t*>(ob + (size_t)j * 1024) = tc::pack_bf16(o[j][0] * inv, o[j][1] * inv);
    }
  }
}

inline uint32_t attn_tc_smem(int HDP) {
  return (uint32_t)(4 + ATT_STAGES * 2) * HDP * 128 * 2 + 4 * 128 * 64 * 2 + 512;
}

// --------------------------------------------------------------------------------------------
// qkv_tc_kernel: RMSGroupNorm -> q|k|v projection (tcgen05) -> RoPE + softmax pre-scale -> bf16 tile images
// (models/mss_tflocoformer.py:452-453, :542-559).  M tiles are sequence-aligned (seq s, tile jt) so each CTA
// writes complete 128-row images (rows >= L come out as exact zeros because their A rows are zero).
//   warp 0: weight image (bulk copy, once)   warp 1: MMA thread   warps 2-5: norm producers
//   warps 6-13: two epilogue groups, each owning half of the columns of every part
// TMEM: three part accumulators (q, k, v) of NPART = heads * HDP columns used as a ring across tiles.
// --------------------------------------------------------------------------------------------
// MUFU approximations (~1 ulp): the operand they scale is rounded to bf16 (2^-9) right after
__device__ __forceinline__ float fast_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

struct QkvTcParams {
  const float* x; SeqMap map; const float* gamma; float eps;
  const char* wimg;                 // [3 parts][C/8 chunks][NPART rows][8] bf16
  const float2* rope;               // [HDP/2][rope_stride] (cos, sin), frequency-major, or nullptr ("nope")
  int rope_stride;                  // positions per table row (L for a per-call table, ROPE_TAB_LEN for the packed one)
  __nv_bfloat16* qkv;               // tile images
  int C, G, L, NTL, nseq, heads, hd, HDP, NPART;
  int n_tiles;
  float qscale;
  int pad_to;                       // image rows are produced up to ceil_pad_to(L): 16 (attn_tc2 / tail kernels) or 64 (attn_tc_kernel)
};

__global__ void rope_table_kernel(float2* __restrict__ tab, const float* __restrict__ freqs, int L, int half, int half_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L * half_pad) return;
  const int j = i / half_pad, f = i - j * half_pad;
  float sn = 0.f, cs = 1.f;
  if (f < half) sincosf((float)j * freqs[f], &sn, &cs);   // same fp32 product as the reference's einsum
  tab[(size_t)f * L + j] = make_float2(cs, sn);           // frequency-major: the 32 rows of a warp read contiguously
}

// The q|k|v weight image carries the RMSGroupNorm gamma (column c of every row is multiplied by gamma[c]: the
// producers of qkv_tc_kernel then only scale by 1 / rms) and, in the q part, the softmax pre-scale log2(e) / sqrt(hd)
// (RoPE is a rotation, so scaling commutes with it): two multiplies per element less in the kernel's hot loops.
__global__ void tc_pack_qkv_kernel(const float* __restrict__ wqkv, const float* __restrict__ wo,
                                   const float* __restrict__ gamma, __nv_bfloat16* __restrict__ qimg,
                                   __nv_bfloat16* __restrict__ oimg, int C, int A, int heads, int hd, int HDP, float qscale) {
  const int NPART = heads * HDP;
  const int n_q = 3 * C * NPART, n_o = NPART * C;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_q + n_o; idx += gridDim.x * blockDim.x) {
    if (idx < n_q) {   // part image: [C/8][NPART][8]; row n = head*HDP + d  <-  wqkv[part*A + head*hd + d][c]
      const int part = idx / (C * NPART), e = idx % (C * NPART);
      const int chunk = e / (NPART * 8), n = (e / 8) % NPART, c = chunk * 8 + (e & 7);
      const int head = n / HDP, d = n % HDP;
      const float w = d < hd ? wqkv[((size_t)part * A + head * hd + d) * C + c] * gamma[c] * (part == 0 ? qscale : 1.f) : 0.f;
      qimg[idx] = __float2bfloat16_rn(w);
    } else {           // wo image: [NPART/8][C rows][8]; row n = out channel, k = head*HDP + d  <-  wo[n][head*hd + d]
      const int e = idx - n_q;
      const int chunk = e / (C * 8), n = (e / 8) % C, k = chunk * 8 + (e & 7);
      const int head = k / HDP, d = k % HDP;
      oimg[e] = __float2bfloat16_rn(d < hd ? wo[(size_t)n * A + head * hd + d] : 0.f);
    }
  }
}

#ifndef QKV_L2_PREFETCH
#define QKV_L2_PREFETCH 1
#endif
constexpr int QKV_PF_DIST = 2;
// (r02, measured with ncu: five warpgroups + setmaxnreg + 32-column tcgen05.ld in the epilogue did not help on the
// frequency axis, 0.669 vs 0.662 ms, and lost on the time axis, 0.887 vs 0.780 ms -- reverted.)
constexpr int QKV_PRODUCER_WARPS = 8;
#ifndef QKV_UF
#define QKV_UF 8   // (8-row, group) units a producer lane keeps in flight (4: two round trips per tile, r01)
#endif
constexpr int QKV_THREADS = 32 * (2 + QKV_PRODUCER_WARPS + 8);

__global__ void __launch_bounds__(QKV_THREADS, 1) qkv_tc_kernel(QkvTcParams p) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.C, NPART = p.NPART, HDP = p.HDP, NTL = p.NTL;
  const uint32_t part_bytes = (uint32_t)C * NPART * 2, a_bytes = (uint32_t)C * 128 * 2;
  const uint32_t off_w = 0, off_a = 3 * part_bytes, off_tab = off_a + 3 * a_bytes, off_bar = off_tab + C * 4;
  const uint32_t sbase = smem_u32(smem);
  auto BAR = [&](int i) { return sbase + off_bar + 8u * i; };
  const int W_FULL = 0, A_FULL = 1, A_EMPTY = 4, D_FULL = 7, D_EMPTY = 10;   // A: 3 slots, D: 3 parts
  constexpr int NPROD = QKV_PRODUCER_WARPS * 32;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + off_bar + 8 * 16);
  if (threadIdx.x == 0) {
    *reinterpret_cast<volatile int*>(smem + off_bar + 8 * 18) = 0;
    mbar_init(BAR(W_FULL), 1);
    for (int i = 0; i < 3; ++i) {
      mbar_init(BAR(A_FULL + i), NPROD); mbar_init(BAR(A_EMPTY + i), 1);
      mbar_init(BAR(D_FULL + i), 1); mbar_init(BAR(D_EMPTY + i), 256);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n_iter = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  griddep_wait();                                          // (PDL) the prologue overlapped the previous kernel's tail

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(BAR(W_FULL), 3 * part_bytes);
      for (int i = 0; i < 3; ++i) bulk_g2s(sbase + off_w + i * part_bytes, p.wimg + (size_t)i * part_bytes, part_bytes, BAR(W_FULL));
    }
#if QKV_L2_PREFETCH
    // L2 prefetch of the x rows of the tiles ahead: the producers keep a whole tile (64 KB) in flight, but only while
    // they load -- not while they normalise -- so HBM latency capped the kernel at ~4 TB/s.  This warp follows the
    // producers (it observes A_FULL) and requests the rows of the tile QKV_PF_DIST tiles further on.
    auto prefetch_tile = [&](int it) {
      if (it >= n_iter) return;
      const int tile = blockIdx.x + it * gridDim.x;
      const int s = tile / NTL, jt = tile - s * NTL;
      const int rows = min(128, p.L - jt * 128);
      const float* src = p.x + p.map.base(s) + (long long)(jt * 128) * p.map.pos_stride;
      if (p.map.pos_stride == C) {                 // rows are contiguous: a few large requests
        const uint32_t total = (uint32_t)rows * C * 4;
        for (uint32_t off = lane * 8192u; off < total; off += 32 * 8192u)
          bulk_prefetch_l2(reinterpret_cast<const char*>(src) + off, min(8192u, total - off));
      } else {
        for (int r = lane; r < rows; r += 32) bulk_prefetch_l2(src + (long long)r * p.map.pos_stride, (uint32_t)C * 4);
      }
    };
    // Pacing by a progress word the producers publish (not by a barrier: a prefetcher that falls a ring turn behind
    // must skip ahead, which a parity wait cannot express).
    volatile int* progress = reinterpret_cast<volatile int*>(smem + off_bar + 8 * 18);
    int next = 1;                                   // next tile (local index) to request
    for (;;) {
      const int done = *progress;                   // tiles the producers have finished
      const int want = min(n_iter, done + 1 + QKV_PF_DIST);
      if (next < done + 1) next = done + 1;         // fell behind: those rows are being read already
      for (; next < want; ++next) prefetch_tile(next);
      if (done >= n_iter - 1 || next >= n_iter) break;
      __nanosleep(200);
    }
#endif
  } else if (warp == 1) {
    {
      const uint32_t idesc = instr_desc(128, NPART);
      const uint32_t hi = (128u >> 4) | (1u << 14);
      const uint32_t lo_a = 128u << 16, lo_b = (uint32_t)NPART << 16;
      const uint32_t a16 = (sbase + off_a) >> 4, w16 = (sbase + off_w) >> 4, as16 = a_bytes >> 4, ps16 = part_bytes >> 4;
      mbar_wait(BAR(W_FULL), 0);
      uint32_t slot = 0, aph = 0;
      for (int it = 0; it < n_iter; ++it) {
        mbar_wait(BAR(A_FULL + slot), aph);
        for (int part = 0; part < 3; ++part) {
          mbar_wait(BAR(D_EMPTY + part), (uint32_t)((it & 1) ^ 1));
          tc_fence_after();
          if (elect_one()) {
            for (int kk = 0; kk < C / 16; ++kk)
              mma_lohi(tmem + part * NPART, (a16 + slot * as16 + kk * 2 * 128) | lo_a, hi,
                       (w16 + part * ps16 + kk * 2 * NPART) | lo_b, hi, idesc, (uint32_t)kk);
            mma_commit(BAR(D_FULL + part));
          }
          __syncwarp();
        }
        if (elect_one()) mma_commit(BAR(A_EMPTY + slot));
        __syncwarp();
        if (++slot == 3) { slot = 0; aph ^= 1; }
      }
    }
  } else if (warp < 2 + QKV_PRODUCER_WARPS) {
    // ---- producers: one (row, group) item per thread and pass; the row stays in registers between the
    //      sum-of-squares and the scaling, so x is read exactly once ----
    const int tp = threadIdx.x - 64;
    const int G = p.G, D = C / G;
    const float rs = rsqrtf((float)D);
    uint32_t slot = 0, ph = 0;
    for (int it = 0; it < n_iter; ++it) {
      mbar_wait(BAR(A_EMPTY + slot), ph ^ 1);
      uint8_t* at = smem + off_a + (size_t)slot * a_bytes;
      const int tile = blockIdx.x + it * gridDim.x;
      const int s = tile / NTL, jt = tile - s * NTL;
      const long long base = p.map.base(s);
      if (D == 32) {
        // Four lanes per (row, group) item, 32 contiguous bytes each: a warp instruction reads 8 rows x 128 B (8 L1
        // tag lookups instead of 32 -- ncu had l1tex throughput at 96 % with one lane per item) and writes one
        // 16-byte chunk per lane, 8 consecutive rows per chunk column: conflict-free.  The group's sum of squares is
        // two shuffles.  Eight items per thread and tile, four in flight at a time.
        const int r8 = lane >> 2, q4 = lane & 3;
        const int pw = warp - 2;                              // producer warp 0 .. QKV_PRODUCER_WARPS - 1
        // rows at or beyond ceil_pad_to(L) are never written to the images (see the epilogue): their A rows are left alone
        const int rows_used = min(128, (p.L + p.pad_to - 1) / p.pad_to * p.pad_to - jt * 128);
        const int n_blk = ((rows_used + 7) >> 3) * G;          // (8-row block, group) units per tile
        // QKV_UF units (2 x 128-bit loads each) in flight per lane: the producers are bound by global-load latency
        // (ncu: the epilogue warps wait on D_FULL, the MMA warp on A_FULL), so a whole tile -- 64 units for 4 groups --
        // is requested in ONE round trip of the 8 producer warps instead of two.
#pragma unroll 1
        for (int u0 = pw; u0 < n_blk; u0 += QKV_UF * QKV_PRODUCER_WARPS) {
          float4 v[QKV_UF][2];
          int rowi[QKV_UF], grpi[QKV_UF];
#pragma unroll
          for (int u = 0; u < QKV_UF; ++u) {
            const int unit = u0 + u * QKV_PRODUCER_WARPS;
            const int rb = unit / G;
            grpi[u] = unit - rb * G;
            rowi[u] = rb * 8 + r8;
            const int j = jt * 128 + rowi[u];
            const bool valid = unit < n_blk && j < p.L;
            const float* src = p.x + base + (long long)j * p.map.pos_stride + grpi[u] * 32 + q4 * 8;
            v[u][0] = valid ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[u][1] = valid ? __ldg(reinterpret_cast<const float4*>(src + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < QKV_UF; ++u) {
            if (u0 + u * QKV_PRODUCER_WARPS >= n_blk) continue;   // warp-uniform
            float ss = v[u][0].x * v[u][0].x + v[u][0].y * v[u][0].y + v[u][0].z * v[u][0].z + v[u][0].w * v[u][0].w +
                       v[u][1].x * v[u][1].x + v[u][1].y * v[u][1].y + v[u][1].z * v[u][1].z + v[u][1].w * v[u][1].w;
            ss += __shfl_xor_sync(0xffffffffu, ss, 1);
            ss += __shfl_xor_sync(0xffffffffu, ss, 2);
            const float inv = fast_rcp(fast_sqrt(ss) * rs + p.eps);
            const int c0 = grpi[u] * 32 + q4 * 8;
            uint4 pk;
            pk.x = pack_bf16(v[u][0].x * inv, v[u][0].y * inv);
            pk.y = pack_bf16(v[u][0].z * inv, v[u][0].w * inv);
            pk.z = pack_bf16(v[u][1].x * inv, v[u][1].y * inv);
            pk.w = pack_bf16(v[u][1].z * inv, v[u][1].w * inv);
            *reinterpret_cast<uint4*>(at + ((size_t)(c0 >> 3) * 128 + rowi[u]) * 16) = pk;
          }
        }
      } else if (D <= 32) {
        // two (row, group) items per thread and pass: 16 independent 128-bit loads in flight, x read exactly once
        for (int item0 = tp; item0 < 128 * G; item0 += 2 * NPROD) {
          float4 v[2][8];
          int rowi[2], grpi[2];
          bool live[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int item = item0 + u * NPROD;
            live[u] = item < 128 * G;
            rowi[u] = item / G; grpi[u] = item - rowi[u] * G;
            const int j = jt * 128 + rowi[u];
            const bool valid = live[u] && j < p.L;
            const float* src = p.x + base + (long long)j * p.map.pos_stride + grpi[u] * D;
#pragma unroll
            for (int d = 0; d < 8; ++d)
              v[u][d] = (valid && 4 * d < D) ? __ldg(reinterpret_cast<const float4*>(src + 4 * d)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (!live[u]) continue;
            float ss = 0.f;
#pragma unroll
            for (int d = 0; d < 8; ++d) ss += v[u][d].x * v[u][d].x + v[u][d].y * v[u][d].y + v[u][d].z * v[u][d].z + v[u][d].w * v[u][d].w;
            const float inv = fast_rcp(fast_sqrt(ss) * rs + p.eps);
#pragma unroll
            for (int d = 0; d < 8; ++d) {
              if (4 * d < D) {
                const int c0 = grpi[u] * D + 4 * d;
                uint2 pk;
                pk.x = pack_bf16(v[u][d].x * inv, v[u][d].y * inv);
                pk.y = pack_bf16(v[u][d].z * inv, v[u][d].w * inv);
                *reinterpret_cast<uint2*>(at + ((size_t)(c0 >> 3) * 128 + rowi[u]) * 16 + (c0 & 7) * 2) = pk;
              }
            }
          }
        }
      } else {
        for (int item = tp; item < 128 * G; item += NPROD) {
          const int row = item / G, grp = item - row * G;
          const int j = jt * 128 + row;
          const bool valid = j < p.L;
          const float* src = p.x + base + (long long)j * p.map.pos_stride + grp * D;
          float ss = 0.f;
          if (valid)
            for (int d = 0; d < D; d += 4) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(src + d));
              ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            }
          const float inv = fast_rcp(fast_sqrt(ss) * rs + p.eps);
          for (int d = 0; d < D; d += 4) {
            uint2 pk = make_uint2(0u, 0u);
            const int c0 = grp * D + d;
            if (valid) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(src + d));
              pk.x = pack_bf16(v.x * inv, v.y * inv);
              pk.y = pack_bf16(v.z * inv, v.w * inv);
            }
            *reinterpret_cast<uint2*>(at + ((size_t)(c0 >> 3) * 128 + row) * 16 + (c0 & 7) * 2) = pk;
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(BAR(A_FULL + slot));
      if (tp == 0) *reinterpret_cast<volatile int*>(smem + off_bar + 8 * 18) = it + 1;   // progress word of the L2 prefetcher
      if (++slot == 3) { slot = 0; ph ^= 1; }
    }
  } else {
    const int ew = warp - (2 + QKV_PRODUCER_WARPS);
    const int e = ew >> 2;                     // column half owned by this group
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const int col_lo = e * (NPART / 2), col_hi = col_lo + NPART / 2;
    const int halfp = HDP / 2;
    const size_t tile_elems = (size_t)HDP * 128;
    for (int it = 0; it < n_iter; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int s = tile / NTL, jt = tile - s * NTL;
      const int j = min(jt * 128 + m, p.L - 1);
      // RoPE factors of this row: shared by q and k and by all heads -> registers, loaded once per tile
      float2 cs[16];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        cs[i] = (p.rope != nullptr && i < halfp) ? __ldg(&p.rope[(size_t)i * p.rope_stride + j]) : make_float2(1.f, 0.f);
      // Rows at or beyond ceil_pad_to(L) are never read by the attention kernels (attn_tc_kernel masks whole 64-key
      // units, attn_tc2_kernel and attn_tail_mma_kernel end on a ceil16 block, attn_tail_rows_kernel reads keys < L):
      // neither loaded from TMEM nor written -- on the time axis (259 rows in three 128-row tiles) that is 29 % of the image.
      const int row_lim = (p.L + p.pad_to - 1) / p.pad_to * p.pad_to - jt * 128;   // rows of this tile anyone reads
      const bool warp_live = quarter * 32 < row_lim;
      for (int part = 0; part < 3; ++part) {
        mbar_wait(BAR(D_FULL + part), (uint32_t)(it & 1));
        tc_fence_after();
        for (int c0 = col_lo; warp_live && c0 < col_hi; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(lane_addr + part * NPART + c0, r);
          tc_wait_ld();
          const int head = c0 / HDP, d0 = c0 - head * HDP;
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            float a = __uint_as_float(r[i]), b = __uint_as_float(r[i + 1]);
            if (part < 2) {
              const float2 f = d0 == 0 ? cs[i >> 1] : cs[8 + (i >> 1)];   // d0 is 0 or 16 (HDP <= 32)
              const float ra = a * f.x - b * f.y, rb = b * f.x + a * f.y;
              a = ra; b = rb;
            }
            w[i >> 1] = pack_bf16(a, b);
          }
          __nv_bfloat16* dst = p.qkv + ((((size_t)part * p.nseq + s) * p.heads + head) * NTL + jt) * tile_elems +
                               ((size_t)(d0 >> 3) * 128 + m) * 8;
          if (m < row_lim) {
            *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(dst + 1024) = make_uint4(w[4], w[5], w[6], w[7]);
          }
        }
        tc_fence_before();
        mbar_arrive(BAR(D_EMPTY + part));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// --------------------------------------------------------------------------------------------
// proj_tc_kernel: x += o . Wo^T (head merge + residual, :536-540, :456).  A tiles are the attention kernel's o
// images (one bulk copy each), Wo stays resident in smem, accumulators double-buffered in TMEM.
//   warp 0: loader   warp 1: MMA thread   warps 2-9: two epilogue groups alternating tiles
// Epilogue (C % 64 == 0): TMEM hands every thread one ROW of the tile, but a warp instruction that touches 32 rows
// of x costs 32 L1 tag cycles -- 8 k cycles per tile, which was the kernel's whole run time.  So the update goes
// through a padded staging tile in shared memory, 64 columns at a time: row-per-thread in, then each warp walks
// its 32 rows with half a warp per row (256 contiguous bytes): the read-modify-write of x is 4 lines per
// instruction instead of 32.
// --------------------------------------------------------------------------------------------
struct ProjTcParams {
  const __nv_bfloat16* oimg; const char* wimg; float* x; SeqMap map;
  int C, AP, L, NTL, n_tiles;
};
constexpr int PROJ_STAGES = 3;
constexpr uint32_t PROJ_STG_PITCH = 64 * 4 + 16;                // staging row: 64 fp32 + 16 B (conflict-free both ways)
constexpr uint32_t PROJ_STG_BYTES = 128 * PROJ_STG_PITCH;

__global__ void __launch_bounds__(320, 1) proj_tc_kernel(ProjTcParams p) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.C, AP = p.AP, NTL = p.NTL;
  const uint32_t w_bytes = (uint32_t)AP * C * 2, a_bytes = (uint32_t)AP * 128 * 2;
  const uint32_t off_w = 0, off_a = w_bytes, off_stg = off_a + PROJ_STAGES * a_bytes, off_bar = off_stg + 2 * PROJ_STG_BYTES;
  const uint32_t sbase = smem_u32(smem);
  auto BAR = [&](int i) { return sbase + off_bar + 8u * i; };
  const int W_FULL = 0, A_FULL = 1, A_EMPTY = 5, D_FULL = 9, D_EMPTY = 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + off_bar + 8 * 16);
  if (threadIdx.x == 0) {
    mbar_init(BAR(W_FULL), 1);
    for (int i = 0; i < PROJ_STAGES; ++i) { mbar_init(BAR(A_FULL + i), 1); mbar_init(BAR(A_EMPTY + i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(D_FULL + i), 1); mbar_init(BAR(D_EMPTY + i), 128); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n_iter = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  griddep_wait();                                          // (PDL) the prologue overlapped the previous kernel's tail

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(BAR(W_FULL), w_bytes);
      bulk_g2s(sbase + off_w, p.wimg, w_bytes, BAR(W_FULL));
    }
    {
      uint32_t slot = 0, ph = 0;
      for (int it = 0; it < n_iter; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        mbar_wait(BAR(A_EMPTY + slot), ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(BAR(A_FULL + slot), a_bytes);
          bulk_g2s(sbase + off_a + slot * a_bytes, p.oimg + (size_t)tile * AP * 128, a_bytes, BAR(A_FULL + slot));
        }
        __syncwarp();
        if (++slot == PROJ_STAGES) { slot = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc = instr_desc(128, C);
      const uint32_t hi = (128u >> 4) | (1u << 14);
      const uint32_t lo_a = 128u << 16, lo_b = (uint32_t)C << 16;
      const uint32_t a16 = (sbase + off_a) >> 4, w16 = (sbase + off_w) >> 4, as16 = a_bytes >> 4;
      mbar_wait(BAR(W_FULL), 0);
      uint32_t slot = 0, ph = 0;
      for (int it = 0; it < n_iter; ++it) {
        const int buf = it & 1;
        mbar_wait(BAR(A_FULL + slot), ph);
        mbar_wait(BAR(D_EMPTY + buf), (uint32_t)(((it >> 1) & 1) ^ 1));
        tc_fence_after();
        if (elect_one()) {
          for (int kk = 0; kk < AP / 16; ++kk)
            mma_lohi(tmem + buf * C, (a16 + slot * as16 + kk * 2 * 128) | lo_a, hi, (w16 + kk * 2 * C) | lo_b, hi, idesc,
                     (uint32_t)kk);
          mma_commit(BAR(D_FULL + buf));
          mma_commit(BAR(A_EMPTY + slot));
        }
        __syncwarp();
        if (++slot == PROJ_STAGES) { slot = 0; ph ^= 1; }
      }
    }
  } else {
    const int e = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    for (int it = e; it < n_iter; it += 2) {
      const int tile = blockIdx.x + it * gridDim.x;
      const int s = tile / NTL, jt = tile - s * NTL;
      const int j = jt * 128 + m;
      mbar_wait(BAR(D_FULL + e), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      float* dst = j < p.L ? p.x + p.map.base(s) + (long long)j * p.map.pos_stride : nullptr;
      int c0 = 0;
      if (C % 64 == 0) {
        uint8_t* stg = smem + off_stg + (size_t)e * PROJ_STG_BYTES;
        const int wq = warp & 3, sub = lane >> 4, l16 = lane & 15;
        float* xb = p.x + p.map.base(s);
        for (; c0 < C; c0 += 64) {
          {
            uint32_t r[64];
            tmem_ld32(lane_addr + e * C + c0, r);
            tmem_ld32(lane_addr + e * C + c0 + 32, r + 32);
            tc_wait_ld();
            uint8_t* row = stg + (size_t)m * PROJ_STG_PITCH;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              *reinterpret_cast<uint4*>(row + 16 * i) = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + e) : "memory");
          // rows wq*32 .. wq*32+31 of the tile, two per instruction, eight instructions in flight
#pragma unroll 1
          for (int i0 = 0; i0 < 16; i0 += 8) {
            float4 xv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int jr = jt * 128 + wq * 32 + 2 * (i0 + i) + sub;
              xv[i] = jr < p.L ? *reinterpret_cast<const float4*>(xb + (long long)jr * p.map.pos_stride + c0 + 4 * l16)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = wq * 32 + 2 * (i0 + i) + sub, jr = jt * 128 + rr;
              const float4 dv = *reinterpret_cast<const float4*>(stg + (size_t)rr * PROJ_STG_PITCH + 16 * l16);
              xv[i].x += dv.x; xv[i].y += dv.y; xv[i].z += dv.z; xv[i].w += dv.w;
              if (jr < p.L) *reinterpret_cast<float4*>(xb + (long long)jr * p.map.pos_stride + c0 + 4 * l16) = xv[i];
            }
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + e) : "memory");
        }
      }
      for (; c0 + 64 <= C; c0 += 64) {     // generic widths: 64 columns per step, row per thread
        uint32_t r[64];
        tmem_ld32(lane_addr + e * C + c0, r);
        tmem_ld32(lane_addr + e * C + c0 + 32, r + 32);
        float4 xv[16];
        if (dst != nullptr) {
#pragma unroll
          for (int i = 0; i < 16; ++i) xv[i] = *reinterpret_cast<const float4*>(dst + c0 + 4 * i);
        }
        tc_wait_ld();
        if (dst != nullptr) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            xv[i].x += __uint_as_float(r[4 * i]); xv[i].y += __uint_as_float(r[4 * i + 1]);
            xv[i].z += __uint_as_float(r[4 * i + 2]); xv[i].w += __uint_as_float(r[4 * i + 3]);
            *reinterpret_cast<float4*>(dst + c0 + 4 * i) = xv[i];
          }
        }
      }
      for (; c0 < C; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(lane_addr + e * C + c0, r);
        tc_wait_ld();
        if (dst != nullptr) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 xv = *reinterpret_cast<const float4*>(dst + c0 + i);
            xv.x += __uint_as_float(r[i]); xv.y += __uint_as_float(r[i + 1]);
            xv.z += __uint_as_float(r[i + 2]); xv.w += __uint_as_float(r[i + 3]);
            *reinterpret_cast<float4*>(dst + c0 + i) = xv;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(BAR(D_EMPTY + e));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

inline bool attn_tc_supported(int C, int heads, int hd) {
  if (hd % 2 != 0 || hd > 32 || C % 16 != 0 || C > 256) return false;   // hd is zero-padded to HDP = 16 or 32 (12 -> 16, 24 -> 32)
  const int HDP = (hd + 15) / 16 * 16, NPART = heads * HDP;
  if (NPART % 32 != 0 || NPART > 128) return false;
  return (size_t)3 * C * NPART * 2 + (size_t)3 * C * 256 + C * 4 + 256 <= (size_t)232448;
}
inline size_t tc_qkv_image_bytes(int C, int heads, int hd) {
  return attn_tc_supported(C, heads, hd) ? (size_t)3 * C * heads * ((hd + 15) / 16 * 16) * 2 : 0;
}
inline size_t tc_wo_image_bytes(int C, int heads, int hd) {
  return attn_tc_supported(C, heads, hd) ? (size_t)C * heads * ((hd + 15) / 16 * 16) * 2 : 0;
}

}  // namespace tfl
