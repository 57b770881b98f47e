// tcgen05 / TMEM kernels of TFL_PRECISION_BF16.
//
// K4  ffn_tc_kernel: x += ConvSwiGLU(RMSGroupNorm(x)) along one axis as ONE persistent,
//     warp-specialised kernel (models/mss_tflocoformer.py:443-447,459-462 -> :626-655):
//       producers   fp32 residual rows -> RMSGroupNorm -> bf16 chunk-major A tile in smem
//       loader      weight stages (pre-packed bf16 smem images) via 1-D bulk async copies
//       MMA thread  conv1d as KT row-shifted tcgen05.mma taps into TMEM (value | gate halves),
//                   transposed conv as KT taps over the SwiGLU'd hidden tile
//       epilogue    TMEM -> bias + SwiGLU -> bf16 hidden tile in smem (never touches HBM);
//                   final: TMEM -> + bias + residual -> x (fp32)
//     Sequences are laid on a "stream" with period P = S + KT - 1 rows (KT - 1 shared zero rows
//     between consecutive sequences == the reference's zero padding AFTER the norm, :640-644),
//     so M tiles run across sequence boundaries and both axes use the same kernel through SeqMap.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace tfl {

constexpr int TC_HC = 32;             // hidden channels per chunk (D1 tile = 2*HC TMEM columns)
constexpr int TC_SMEM_MAX = 232448;   // 227 KB opt-in shared memory per CTA

struct FfnTcGeom {
  int C, H, KT, G, NT, NS;            // NS = weight ring stages
  int AR;                             // rows per A / G tile = 128 + KT - 1
  int TS;                             // output rows per tile = 128 - (KT - 1)
  int NC, KS;                         // hidden chunks, W2 stages per chunk
  uint32_t stage_bytes, a_slot_bytes, g_buf_bytes;
  uint32_t off_a, off_g, off_w, off_tab, off_bar, smem_bytes;
  int threads;
};

inline bool ffn_tc_geometry(int C, int H, int KT, int G, FfnTcGeom* g) {
  if (C % 16 != 0 || C > 256 || H % TC_HC != 0 || (C / G) % 4 != 0 || KT < 1 || KT > 8) return false;
  g->C = C; g->H = H; g->KT = KT; g->G = G;
  g->NT = C <= 128 ? 2 : 1;
  g->AR = 128 + KT - 1; g->TS = 128 - (KT - 1);
  g->NC = H / TC_HC; g->KS = (KT + 1) / 2;
  g->stage_bytes = 128u * C;
  g->a_slot_bytes = (uint32_t)(C / 8) * g->AR * 16;
  g->g_buf_bytes = (uint32_t)(TC_HC / 8) * g->AR * 16;
  uint32_t off = 0;
  g->off_a = off; off += (g->NT + 1) * g->a_slot_bytes;
  g->off_g = off; off += g->NT * 2 * g->g_buf_bytes;
  g->off_tab = off; off += (2 * H + 2 * C) * 4;
  off = (off + 15) & ~15u;
  g->off_bar = off; off += 512;
  off = (off + 127) & ~127u;
  g->off_w = off;
  if (off + 2 * g->stage_bytes > (uint32_t)TC_SMEM_MAX) return false;
  g->NS = (int)((TC_SMEM_MAX - off) / g->stage_bytes);
  if (g->NS > 8) g->NS = 8;
  g->smem_bytes = off + g->NS * g->stage_bytes;
  g->threads = 32 * (6 + 4 * g->NT);
  return true;
}

inline size_t tc_ffn_image_bytes(int C, int H, int K) {
  if (C % 16 != 0 || C > 256 || H % TC_HC != 0 || K < 1 || K > 8) return 0;
  return (size_t)(H / TC_HC) * (K + (K + 1) / 2) * 128 * C;
}

// Weight image in consumption order: per hidden chunk c the KT conv1d tap stages W1(c, k)
// (B operand [64 rows = 32 value | 32 gate] x [C], chunk-major), then -- one chunk late, matching
// the MMA issue order -- the KS transposed-conv stages W2(c - 1, pair) (2 taps x [C rows] x [32]).
__global__ void tc_pack_ffn_kernel(const float* __restrict__ w1, const float* __restrict__ w2,
                                   __nv_bfloat16* __restrict__ img, int C, int H, int KT) {
  const int NC = H / TC_HC, KS = (KT + 1) / 2, n_per = KT + KS;
  const long long stage_elems = 64LL * C;
  const long long total = (long long)NC * n_per * stage_elems;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(idx / stage_elems);
    const int e = (int)(idx % stage_elems);
    int is_w1, c, sub;
    if (g < KT) { is_w1 = 1; c = 0; sub = g; }
    else {
      const int gp = g - KT, blk = gp / n_per, rem = gp % n_per;
      if (blk < NC - 1) {
        if (rem < KT) { is_w1 = 1; c = blk + 1; sub = rem; } else { is_w1 = 0; c = blk; sub = rem - KT; }
      } else { is_w1 = 0; c = NC - 1; sub = rem; }
    }
    float v = 0.f;
    if (is_w1) {
      const int chunk = e / (64 * 8), n = (e / 8) % 64, cc = chunk * 8 + (e & 7);
      const int row = n < TC_HC ? c * TC_HC + n : H + c * TC_HC + (n - TC_HC);
      v = w1[((size_t)row * C + cc) * KT + sub];
    } else {
      const int per_tap = 4 * C * 8;
      const int tl = e / per_tap, e2 = e % per_tap;
      const int chunk = e2 / (C * 8), n = (e2 / 8) % C, hh = chunk * 8 + (e2 & 7);
      const int tap = 2 * sub + tl;
      if (tap < KT) v = w2[((size_t)(c * TC_HC + hh) * C + n) * KT + (KT - 1 - tap)];
    }
    img[idx] = __float2bfloat16_rn(v);
  }
}

inline int tc_pack_ffn(const float* w1, const float* b1, const float* w2, const float* b2, char* img, int C, int H,
                       int K, cudaStream_t st) {
  (void)b1; (void)b2;
  if (tc_ffn_image_bytes(C, H, K) == 0) return 0;  // shape not covered by the tcgen05 path
  tc_pack_ffn_kernel<<<592, 256, 0, st>>>(w1, w2, (__nv_bfloat16*)img, C, H, K);
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

__global__ void __launch_bounds__(448, 1) ffn_tc_kernel(FfnTcParams p, FfnTcGeom g) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = g.C, H = g.H, KT = g.KT, NT = g.NT, NC = g.NC, KS = g.KS, AR = g.AR, TS = g.TS, NS = g.NS;
  const int NA = NT + 1;
  const uint32_t sbase = smem_u32(smem);
  float* tab_b1 = reinterpret_cast<float*>(smem + g.off_tab);
  float* tab_b2 = tab_b1 + 2 * H;
  float* tab_gamma = tab_b2 + C;
  // barrier map (8 bytes each)
  const uint32_t bar0 = sbase + g.off_bar;
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  // 0..7 w_full, 8..15 w_empty, 16..18 a_full, 19..21 a_empty, 22..25 d1_full[tile][buf], 26..29 d1_empty,
  // 30..33 g_full, 34..37 g_empty, 38..39 d2_full, 40..41 d2_empty; 48: tmem base slot
  const int W_FULL = 0, W_EMPTY = 8, A_FULL = 16, A_EMPTY = 19, D1_FULL = 22, D1_EMPTY = 26, G_FULL = 30, G_EMPTY = 34,
            D2_FULL = 38, D2_EMPTY = 40;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + g.off_bar + 8 * 48);

  for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) tab_b1[i] = p.b1[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { tab_b2[i] = p.b2[i]; tab_gamma[i] = p.gamma[i]; }
  {  // hidden tiles: rows >= 128 are only ever read for discarded output rows; keep them finite
    uint32_t* gz = reinterpret_cast<uint32_t*>(smem + g.off_g);
    for (uint32_t i = threadIdx.x; i < NT * 2 * g.g_buf_bytes / 4; i += blockDim.x) gz[i] = 0u;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(BAR(W_FULL + i), 1); mbar_init(BAR(W_EMPTY + i), 1); }
    for (int i = 0; i < 3; ++i) { mbar_init(BAR(A_FULL + i), 128); mbar_init(BAR(A_EMPTY + i), 1); }
    for (int i = 0; i < 4; ++i) {
      mbar_init(BAR(D1_FULL + i), 1); mbar_init(BAR(D1_EMPTY + i), 128);
      mbar_init(BAR(G_FULL + i), 128); mbar_init(BAR(G_EMPTY + i), 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(D2_FULL + i), 1); mbar_init(BAR(D2_EMPTY + i), 128); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int n_pairs = (p.n_tiles + NT - 1) / NT;
  const int n_iter = (n_pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int stages_per_iter = NC * (KT + KS);
  const uint32_t d2_col0 = 4u * 2 * TC_HC;  // after the four D1 buffers

  if (warp == 0) {
    // ===================== weight loader =====================
    if (lane == 0) {
      long long n = 0;
      for (int it = 0; it < n_iter; ++it)
        for (int s = 0; s < stages_per_iter; ++s, ++n) {
          const int slot = (int)(n % NS);
          mbar_wait(BAR(W_EMPTY + slot), (uint32_t)(((n / NS) & 1) ^ 1));
          mbar_arrive_expect_tx(BAR(W_FULL + slot), g.stage_bytes);
          bulk_g2s(sbase + g.off_w + slot * g.stage_bytes, p.img + (size_t)s * g.stage_bytes, g.stage_bytes,
                   BAR(W_FULL + slot));
        }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc1 = instr_desc(128, 2 * TC_HC), idesc2 = instr_desc(128, C);
      const uint32_t lbo_a = AR * 16, lbo_b1 = 64 * 16, lbo_g = AR * 16, lbo_b2 = C * 16;
      long long wn = 0;
      auto wait_stage = [&]() -> uint32_t {
        const int slot = (int)(wn % NS);
        mbar_wait(BAR(W_FULL + slot), (uint32_t)((wn / NS) & 1));
        tc_fence_after();
        return sbase + g.off_w + slot * g.stage_bytes;
      };
      auto release_stage = [&]() { mma_commit(BAR(W_EMPTY + (int)(wn % NS))); ++wn; };
      for (int it = 0; it < n_iter; ++it) {
        auto mma2 = [&](int cc) {
          const long long qq = (long long)it * NC + cc;
          const int gbuf = (int)(qq & 1);
          for (int t = 0; t < NT; ++t) mbar_wait(BAR(G_FULL + t * 2 + gbuf), (uint32_t)((qq >> 1) & 1));
          if (cc == 0) for (int t = 0; t < NT; ++t) mbar_wait(BAR(D2_EMPTY + t), (uint32_t)((it & 1) ^ 1));
          tc_fence_after();
          for (int sp = 0; sp < KS; ++sp) {
            const uint32_t wb = wait_stage();
            for (int t = 0; t < NT; ++t) {
              const uint32_t gb = sbase + g.off_g + (t * 2 + gbuf) * g.g_buf_bytes;
              for (int tl = 0; tl < 2; ++tl) {
                const int tap = 2 * sp + tl;
                if (tap >= KT) break;
                for (int kk = 0; kk < TC_HC / 16; ++kk) {
                  const uint64_t ad = smem_desc(gb + tap * 16 + kk * 2 * lbo_g, lbo_g, 128);
                  const uint64_t bd = smem_desc(wb + tl * (4 * C * 16) + kk * 2 * lbo_b2, lbo_b2, 128);
                  mma_ss(tmem + d2_col0 + t * C, ad, bd, idesc2, !(cc == 0 && tap == 0 && kk == 0));
                }
              }
            }
            release_stage();
          }
          for (int t = 0; t < NT; ++t) mma_commit(BAR(G_EMPTY + t * 2 + gbuf));
        };
        for (int c = 0; c < NC; ++c) {
          const long long q = (long long)it * NC + c;
          const int buf = (int)(q & 1);
          for (int t = 0; t < NT; ++t) {
            if (c == 0) {
              const long long n = (long long)it * NT + t;
              mbar_wait(BAR(A_FULL + (int)(n % NA)), (uint32_t)((n / NA) & 1));
            }
            mbar_wait(BAR(D1_EMPTY + t * 2 + buf), (uint32_t)(((q >> 1) & 1) ^ 1));
          }
          tc_fence_after();
          for (int k = 0; k < KT; ++k) {
            const uint32_t wb = wait_stage();
            for (int t = 0; t < NT; ++t) {
              const long long n = (long long)it * NT + t;
              const uint32_t ab = sbase + g.off_a + (uint32_t)(n % NA) * g.a_slot_bytes;
              for (int kk = 0; kk < C / 16; ++kk) {
                const uint64_t ad = smem_desc(ab + k * 16 + kk * 2 * lbo_a, lbo_a, 128);
                const uint64_t bd = smem_desc(wb + kk * 2 * lbo_b1, lbo_b1, 128);
                mma_ss(tmem + (t * 2 + buf) * (2 * TC_HC), ad, bd, idesc1, (k | kk) != 0);
              }
            }
            release_stage();
          }
          for (int t = 0; t < NT; ++t) mma_commit(BAR(D1_FULL + t * 2 + buf));
          if (c == NC - 1)
            for (int t = 0; t < NT; ++t) mma_commit(BAR(A_EMPTY + (int)(((long long)it * NT + t) % NA)));
          if (c > 0) mma2(c - 1);
        }
        mma2(NC - 1);
        for (int t = 0; t < NT; ++t) mma_commit(BAR(D2_FULL + t));
      }
    }
  } else if (warp < 6) {
    // ===================== A producers: x -> RMSGroupNorm -> bf16 chunk-major tile =====================
    const int tp = threadIdx.x - 64;  // 0..127
    const int G = g.G, D = C / G;
    const float rs = rsqrtf((float)D);
    for (int it = 0; it < n_iter; ++it)
      for (int t = 0; t < NT; ++t) {
        const long long n = (long long)it * NT + t;
        const int slot = (int)(n % NA);
        mbar_wait(BAR(A_EMPTY + slot), (uint32_t)(((n / NA) & 1) ^ 1));
        uint8_t* at = smem + g.off_a + (size_t)slot * g.a_slot_bytes;
        const long long tile = ((long long)blockIdx.x + (long long)it * gridDim.x) * NT + t;
        const long long r0 = tile * TS;
        for (int item = tp; item < AR * G; item += 128) {
          const int row = item / G, grp = item - row * G;
          const long long r = r0 + row;
          bool valid = r < p.R;
          const float* src = nullptr;
          if (valid) {
            const int s = (int)(r / p.P), j = (int)(r - (long long)s * p.P);
            valid = j >= KT - 1;
            if (valid) src = p.x + p.map.base(s) + (long long)(j - (KT - 1)) * p.map.pos_stride + grp * D;
          }
          float ss = 0.f;
          if (valid)
            for (int d = 0; d < D; d += 4) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(src + d));
              ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
            }
          const float denom = sqrtf(ss) * rs + p.eps;
          for (int d = 0; d < D; d += 4) {
            uint2 pk = make_uint2(0u, 0u);
            const int c0 = grp * D + d;
            if (valid) {
              const float4 v = __ldg(reinterpret_cast<const float4*>(src + d));  // L1 hit
              pk.x = pack_bf16(v.x / denom * tab_gamma[c0], v.y / denom * tab_gamma[c0 + 1]);
              pk.y = pack_bf16(v.z / denom * tab_gamma[c0 + 2], v.w / denom * tab_gamma[c0 + 3]);
            }
            *reinterpret_cast<uint2*>(at + ((size_t)(c0 >> 3) * AR + row) * 16 + (c0 & 7) * 2) = pk;
          }
        }
        fence_proxy_async();
        mbar_arrive(BAR(A_FULL + slot));
      }
  } else {
    // ===================== epilogue groups (one per tile slot) =====================
    const int t = (warp - 6) >> 2;
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int m = quarter * 32 + lane;       // tile row
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    for (int it = 0; it < n_iter; ++it) {
      for (int c = 0; c < NC; ++c) {
        const long long q = (long long)it * NC + c;
        const int buf = (int)(q & 1);
        const uint32_t par = (uint32_t)((q >> 1) & 1);
        mbar_wait(BAR(D1_FULL + t * 2 + buf), par);
        tc_fence_after();
        uint32_t rv[32], rg[32];
        tmem_ld32(lane_addr + (t * 2 + buf) * (2 * TC_HC), rv);
        tmem_ld32(lane_addr + (t * 2 + buf) * (2 * TC_HC) + TC_HC, rg);
        tc_wait_ld();
        tc_fence_before();
        mbar_arrive(BAR(D1_EMPTY + t * 2 + buf));
        mbar_wait(BAR(G_EMPTY + t * 2 + buf), par ^ 1);
        uint8_t* gt = smem + g.off_g + (size_t)(t * 2 + buf) * g.g_buf_bytes;
        const float* bv = tab_b1 + c * TC_HC;
        const float* bg = tab_b1 + H + c * TC_HC;
#pragma unroll
        for (int ch = 0; ch < TC_HC / 8; ++ch) {
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float hv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int i = ch * 8 + e * 2 + u;
              const float val = __uint_as_float(rv[i]) + bv[i];
              const float gate = __uint_as_float(rg[i]) + bg[i];
              hv[u] = val * gate * __frcp_rn(1.f + __expf(-gate));   // value * SiLU(gate), :648-649
            }
            o[e] = pack_bf16(hv[0], hv[1]);
          }
          *reinterpret_cast<uint4*>(gt + ((size_t)ch * AR + m) * 16) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        fence_proxy_async();
        mbar_arrive(BAR(G_FULL + t * 2 + buf));
      }
      // ---- final: transposed-conv accumulator + bias + residual -> x ----
      mbar_wait(BAR(D2_FULL + t), (uint32_t)(it & 1));
      tc_fence_after();
      const long long tile = ((long long)blockIdx.x + (long long)it * gridDim.x) * NT + t;
      const long long ro = tile * TS + m;
      float* dst = nullptr;
      const float* res = nullptr;
      if (m < TS && ro < p.R) {
        const int s = (int)(ro / p.P), i = (int)(ro - (long long)s * p.P);
        if (i < p.S) {
          const long long off = p.map.base(s) + (long long)i * p.map.pos_stride;
          dst = p.y + off; res = p.x + off;
        }
      }
      for (int c0 = 0; c0 < C; c0 += 16) {
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// --------------------------------------------------------------------------------------------
// Self-test of the tcgen05 plumbing (descriptors, row-shifted taps, bulk copy, TMEM round trip):
// D[128, N] = sum_tap A[m + tap, :] . B_tap[n, :]   (mode 0: B K-major, staged by a bulk copy)
// D[128, N] = A[m, :] . V[:, n]                     (mode 1: V MN-major, rows of V are the K index)
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
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 256);
  for (int i = threadIdx.x; i < AR * Kd; i += blockDim.x) {
    const int r = i / Kd, c = i % Kd;
    *reinterpret_cast<__nv_bfloat16*>(sa + ((size_t)(c >> 3) * AR + r) * 16 + (c & 7) * 2) = __float2bfloat16_rn(A[i]);
  }
  if (mode == 1) {
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
        mma_ss(tmem, ad, bd, idesc, kk != 0);
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
  if (warp == 0) tmem_dealloc(tmem, 256);
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
  TFL_CHECK(mode == 0 || taps == 1, "mode 1 uses a single tap");
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

// y = x + ConvSwiGLU(RMSGroupNorm(x)); x and y must be distinct buffers.
inline int tc_ffn(const tfl_plan* pl, const char* packed, int layer, int axis, int j, const float* x, float* y, int B,
                  int Tf, int F, cudaStream_t st) {
  TFL_CHECK(x != y, "tc_ffn needs distinct input and output buffers");
  const tfl_config& c = pl->cfg;
  const FfnPack& f = pl->lay.paths[(size_t)layer * 2 + axis].ffn[j];
  FfnTcGeom g;
  TFL_CHECK(ffn_tc_geometry(c.emb_dim, f.hidden, c.conv_kernel, c.num_groups, &g),
            "bf16 tcgen05 FFN needs emb_dim %% 16 == 0 (<= 256), ffn_hidden %% 32 == 0, conv1d_kernel <= 8 "
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
  static thread_local uint32_t smem_set = 0;
  if (g.smem_bytes > smem_set) {
    TFL_CUDA(cudaFuncSetAttribute(ffn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
    smem_set = g.smem_bytes;
  }
  const int n_pairs = (p.n_tiles + g.NT - 1) / g.NT;
  const int grid = n_pairs < pl->sm_count ? n_pairs : pl->sm_count;
  ffn_tc_kernel<<<grid, g.threads, g.smem_bytes, st>>>(p, g);
  TFL_LAUNCH_CHECK();
  return 0;
}

}  // namespace tfl
