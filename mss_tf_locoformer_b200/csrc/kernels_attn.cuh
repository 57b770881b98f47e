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
  int nseq, heads, L, NTL, HDP, NP;   // NP = query-tile pairs per (seq, head)
  int n_items;
};

constexpr int ATT_STAGES = 4;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(320, 1) attn_tc_kernel(AttnTcParams p) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HDP = p.HDP, NTL = p.NTL;
  const uint32_t tile_bytes = (uint32_t)HDP * 128 * 2;          // one Q / K / V tile
  const uint32_t p_bytes = 128u * 128 * 2;                       // one P tile
  const uint32_t off_q = 0, off_kv = 2 * tile_bytes, off_p = off_kv + ATT_STAGES * 2 * tile_bytes;
  const uint32_t off_bar = off_p + 2 * p_bytes;
  const uint32_t sbase = smem_u32(smem);
  auto BAR = [&](int i) { return sbase + off_bar + 8u * i; };
  // 0..3 kv_full, 4..7 kv_empty, 8..9 q_full[g], 10..11 q_empty[g], 12..13 s_full[g], 14..15 s_empty[g],
  // 16..17 p_full[g], 18..19 p_empty[g], 20..23 o_full[g][buf], 24..27 o_empty[g][buf]; slot 32: TMEM base
  const int KV_FULL = 0, KV_EMPTY = 4, Q_FULL = 8, Q_EMPTY = 10, S_FULL = 12, S_EMPTY = 14, P_FULL = 16, P_EMPTY = 18,
            O_FULL = 20, O_EMPTY = 24;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + off_bar + 8 * 32);
  if (threadIdx.x == 0) {
    for (int i = 0; i < ATT_STAGES; ++i) { mbar_init(BAR(KV_FULL + i), 1); mbar_init(BAR(KV_EMPTY + i), 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(BAR(Q_FULL + g), 1); mbar_init(BAR(Q_EMPTY + g), 1);
      mbar_init(BAR(S_FULL + g), 1); mbar_init(BAR(S_EMPTY + g), 128);
      mbar_init(BAR(P_FULL + g), 128); mbar_init(BAR(P_EMPTY + g), 1);
      for (int b = 0; b < 2; ++b) { mbar_init(BAR(O_FULL + g * 2 + b), 1); mbar_init(BAR(O_EMPTY + g * 2 + b), 128); }
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t o_col0 = 256;  // S[0] 0..127, S[1] 128..255, O_j[g][buf] at 256 + (g*2+buf)*HDP
  const size_t which_stride = (size_t)p.nseq * p.heads * NTL * HDP * 128;   // elements between q, k, v planes

  if (warp == 0) {
    // ===================== loader =====================
    if (lane == 0) {
      uint32_t kslot = 0, kph = 0, qph[2] = {0, 0};
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int pair = item % p.NP, sh = item / p.NP;            // sh = seq * heads + head
        const __nv_bfloat16* qb = p.qkv + (size_t)sh * NTL * HDP * 128;
        const int nq = min(2, NTL - 2 * pair);
        for (int g = 0; g < nq; ++g) {
          mbar_wait(BAR(Q_EMPTY + g), qph[g] ^ 1);
          qph[g] ^= 1;
          mbar_arrive_expect_tx(BAR(Q_FULL + g), tile_bytes);
          bulk_g2s(sbase + off_q + g * tile_bytes, qb + (size_t)(2 * pair + g) * HDP * 128, tile_bytes, BAR(Q_FULL + g));
        }
        for (int j = 0; j < NTL; ++j) {
          mbar_wait(BAR(KV_EMPTY + kslot), kph ^ 1);
          mbar_arrive_expect_tx(BAR(KV_FULL + kslot), 2 * tile_bytes);
          const uint32_t dst = sbase + off_kv + kslot * 2 * tile_bytes;
          bulk_g2s(dst, qb + which_stride + (size_t)j * HDP * 128, tile_bytes, BAR(KV_FULL + kslot));
          bulk_g2s(dst + tile_bytes, qb + 2 * which_stride + (size_t)j * HDP * 128, tile_bytes, BAR(KV_FULL + kslot));
          if (++kslot == ATT_STAGES) { kslot = 0; kph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_s = instr_desc(128, 128), idesc_pv = instr_desc(128, HDP, /*b_mn_major=*/true);
      const uint32_t hi_k = (128u >> 4) | (1u << 14);              // K-major tiles: SBO = 128 B
      const uint32_t lo_k = 128u << 16;                            //                LBO = 128 rows * 16 B
      const uint32_t hi_v = ((128u * 16) >> 4) | (1u << 14);       // V as MN-major: SBO = 2048 B (next 8 columns)
      const uint32_t lo_v = (128u >> 4) << 16;                     //                LBO = 128 B (next 8 kv rows)
      const uint32_t q16 = (sbase + off_q) >> 4, kv16 = (sbase + off_kv) >> 4, p16 = (sbase + off_p) >> 4;
      const uint32_t tile16 = tile_bytes >> 4, pt16 = p_bytes >> 4;
      uint32_t kslot = 0, kph = 0, qph[2] = {0, 0};
      uint32_t sph[2] = {0, 0}, pph[2] = {0, 0}, oph[2][2] = {{0, 0}, {0, 0}};
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const int pair = item % p.NP;
        const int nq = min(2, NTL - 2 * pair);
        auto issue_s = [&](int g, uint32_t slot) {
          mbar_wait(BAR(S_EMPTY + g), sph[g] ^ 1);
          tc_fence_after();
          const uint32_t qa = q16 + g * tile16, kb = kv16 + slot * 2 * tile16;
          for (int kk = 0; kk < HDP / 16; ++kk)
            mma_lohi(tmem + g * 128, (qa + kk * 2 * 128) | lo_k, hi_k, (kb + kk * 2 * 128) | lo_k, hi_k, idesc_s, (uint32_t)kk);
          mma_commit(BAR(S_FULL + g));
          sph[g] ^= 1;
        };
        for (int g = 0; g < nq; ++g) { mbar_wait(BAR(Q_FULL + g), qph[g]); qph[g] ^= 1; }
        mbar_wait(BAR(KV_FULL + kslot), kph);
        tc_fence_after();
        for (int g = 0; g < nq; ++g) issue_s(g, kslot);
        for (int j = 0; j < NTL; ++j) {
          uint32_t nslot = kslot + 1, nph = kph;
          if (nslot == ATT_STAGES) { nslot = 0; nph ^= 1; }
          const bool more = j + 1 < NTL;
          if (more) { mbar_wait(BAR(KV_FULL + nslot), nph); tc_fence_after(); }
          for (int g = 0; g < nq; ++g) {
            mbar_wait(BAR(P_FULL + g), pph[g]);        // softmax of tile j done (S[g] was released before)
            pph[g] ^= 1;
            if (more) issue_s(g, nslot);
            else if (g == nq - 1) for (int gg = 0; gg < nq; ++gg) mma_commit(BAR(Q_EMPTY + gg));  // all S MMAs issued
            const int ob = j & 1;
            mbar_wait(BAR(O_EMPTY + g * 2 + ob), oph[g][ob] ^ 1);
            oph[g][ob] ^= 1;
            tc_fence_after();
            const uint32_t pa = p16 + g * pt16, vb = kv16 + kslot * 2 * tile16 + tile16;
            for (int kk = 0; kk < 8; ++kk)
              mma_lohi(tmem + o_col0 + (g * 2 + ob) * HDP, (pa + kk * 2 * 128) | lo_k, hi_k, (vb + kk * 16) | lo_v, hi_v,
                       idesc_pv, (uint32_t)kk);
            mma_commit(BAR(O_FULL + g * 2 + ob));
            mma_commit(BAR(P_EMPTY + g));
          }
          mma_commit(BAR(KV_EMPTY + kslot));
          kslot = nslot; kph = nph;
        }
      }
    }
  } else {
    // ===================== softmax groups =====================
    const int g = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    uint8_t* pt = smem + off_p + (size_t)g * p_bytes;
    uint32_t sph = 0, pph = 0, oph[2] = {0, 0};
    const int OC = HDP / 8;                     // 16-byte chunks per output row and head
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const int pair = item % p.NP, sh = item / p.NP;
      const int nq = min(2, NTL - 2 * pair);
      if (g >= nq) continue;
      const int qt = 2 * pair + g;
      float m_run = -INFINITY, l_run = 0.f, alpha = 0.f;
      float o[32];
#pragma unroll
      for (int d = 0; d < 32; ++d) o[d] = 0.f;
      auto fold = [&](int ob) {               // o += O_j of the previous tile
        mbar_wait(BAR(O_FULL + g * 2 + ob), oph[ob]);
        oph[ob] ^= 1;
        tc_fence_after();
        for (int c0 = 0; c0 < HDP; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(lane_addr + o_col0 + (g * 2 + ob) * HDP + c0, r);
          tc_wait_ld();
#pragma unroll
          for (int e = 0; e < 16; ++e) if (c0 + e < 32) o[(c0 + e) & 31] += __uint_as_float(r[e]);
        }
        tc_fence_before();
        mbar_arrive(BAR(O_EMPTY + g * 2 + ob));
      };
      for (int j = 0; j < NTL; ++j) {
        mbar_wait(BAR(S_FULL + g), sph);
        sph ^= 1;
        tc_fence_after();
        uint32_t s[128];
        tmem_ld32(lane_addr + g * 128, s);
        tmem_ld32(lane_addr + g * 128 + 32, s + 32);
        tmem_ld32(lane_addr + g * 128 + 64, s + 64);
        tmem_ld32(lane_addr + g * 128 + 96, s + 96);
        tc_wait_ld();
        tc_fence_before();
        mbar_arrive(BAR(S_EMPTY + g));
        const int valid = min(128, p.L - j * 128);
        float mx = m_run;
#pragma unroll
        for (int i = 0; i < 128; ++i) {
          if (i >= valid) s[i] = 0xff800000u;  // -inf
          mx = fmaxf(mx, __uint_as_float(s[i]));
        }
        alpha = fast_exp2(m_run - mx);
        m_run = mx;
        float rs = 0.f;
        mbar_wait(BAR(P_EMPTY + g), pph ^ 1);   // P.V of the previous tile has consumed the P buffer
        pph ^= 1;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float p0 = fast_exp2(__uint_as_float(s[c * 8 + 2 * e]) - mx);
            const float p1 = fast_exp2(__uint_as_float(s[c * 8 + 2 * e + 1]) - mx);
            rs += p0 + p1;
            w[e] = pack_bf16(p0, p1);
          }
          *reinterpret_cast<uint4*>(pt + ((size_t)c * 128 + m) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_proxy_async();
        mbar_arrive(BAR(P_FULL + g));
        if (j > 0) fold((j - 1) & 1);
        l_run = l_run * alpha + rs;
#pragma unroll
        for (int d = 0; d < 32; ++d) o[d] *= alpha;
      }
      fold((NTL - 1) & 1);
      // ---- normalise and store this head's slice of the o image ----
      const int s_idx = sh / p.heads, h = sh - s_idx * p.heads;
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

inline uint32_t attn_tc_smem(int HDP) {
  return (uint32_t)(2 + ATT_STAGES * 2) * HDP * 128 * 2 + 2 * 128 * 128 * 2 + 512;
}

// ---- interim glue while the projections still run on the fp32 tap-GEMM ------------------------------
struct EpiQkvImg {  // tap-GEMM epilogue: RoPE + softmax pre-scale, bf16 tile images (layout above)
  __nv_bfloat16* img; int A, hd, heads, L, nseq, NTL, HDP; const float* freqs; float qscale;
  __device__ __forceinline__ void operator()(int s, int j, long long r, int n0, int N, const float* v) const {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      const int n = n0 + i;
      if (n >= N) break;
      const int which = n / A, rem = n - which * A, head = rem / hd, d = rem - head * hd;
      float a = v[i], b = v[i + 1];
      if (freqs != nullptr && which < 2) {
        const float ang = (float)j * __ldg(&freqs[d >> 1]);
        float sn, cs;
        sincosf(ang, &sn, &cs);
        const float ra = a * cs - b * sn, rb = b * cs + a * sn;
        a = ra; b = rb;
      }
      if (which == 0) { a *= qscale; b *= qscale; }
      const size_t tile = (((size_t)which * nseq + s) * heads + head) * NTL + (j >> 7);
      __nv_bfloat162 pk = __floats2bfloat162_rn(a, b);
      *reinterpret_cast<__nv_bfloat162*>(img + tile * ((size_t)HDP * 128) + ((size_t)(d >> 3) * 128 + (j & 127)) * 8 + (d & 7)) = pk;
    }
  }
};

__global__ void oimg_to_f32_kernel(const __nv_bfloat16* __restrict__ img, float* __restrict__ o, int nseq, int L,
                                   int NTL, int heads, int hd, int HDP) {
  const int A = heads * hd;
  const long long total = (long long)nseq * L * A;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(i % A);
    const long long row = i / A;
    const int j = (int)(row % L), s = (int)(row / L);
    const int h = a / hd, d = a - h * hd;
    const size_t chunk = (size_t)h * (HDP / 8) + (d >> 3);
    o[i] = __bfloat162float(img[(((size_t)s * NTL + (j >> 7)) * (heads * (HDP / 8)) + chunk) * 1024 + (size_t)(j & 127) * 8 + (d & 7)]);
  }
}

}  // namespace tfl
