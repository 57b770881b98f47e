// K8  Band-split encoder / band-wise decoder of BS-Locoformer
// (standalone/bslocoformer_separator.py:186-270), fp32 CUDA-core kernels.  The reference runs
// 62 x (GroupNorm + 1x1 conv) and 62 x (GroupNorm + 3 x 1x1 conv + tanh + GLU) tiny module calls;
// here each stage is ONE launch with a CTA per (band, frame tile, batch element).
//
// Weights live in one caller-packed fp32 buffer `w`; `table` (device, int64) holds per band b
//   table[b*16 + 0]  = first bin f0          table[b*16 + 1]  = width w_b
//   split : [2] gn_w  [3] gn_b  [4] Wt [K_b][C] (transposed 1x1 conv)  [5] bias [C]      K_b = w_b * 2M
//   decode: [6] gn_w  [7] gn_b  [8] W1t [C][4C]  [9] b1  [10] W3t [4C][4C]  [11] b3  [12] W4t [4C][O_b]  [13] b4
//   O_b = w_b * S * 2M * 2 (pre-GLU).
#pragma once
#include "common.cuh"

namespace tfl {

constexpr int BS_TT = 16;  // frames per CTA tile

__device__ __forceinline__ float2 block_sum2(float a, float b, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  __syncthreads();
  if (lane == 0) { red[warp * 2] = a; red[warp * 2 + 1] = b; }
  __syncthreads();
  float sa = 0.f, sb = 0.f;
  for (int i = 0; i < nw; ++i) { sa += red[i * 2]; sb += red[i * 2 + 1]; }
  return make_float2(sa, sb);
}

// spec [B, M, T, F, 2] -> x [B, T, nb, C].  Band channel index k = f_local * 2M + ch, ch = (re_0..re_{M-1}, im_0..)
// (:163-164, :247-252).  GroupNorm(1, K_b) statistics run over (K_b, T) per batch element.
__global__ void __launch_bounds__(256) bs_split_kernel(const float* __restrict__ spec, int M, int T, int F, int C, int nb,
                                                       const long long* __restrict__ table, const float* __restrict__ w,
                                                       float* __restrict__ x, float eps) {
  extern __shared__ float sm[];           // xn [K_b][BS_TT] + reduction scratch
  const int band = blockIdx.x, b = blockIdx.z;
  const long long* tb = table + band * 16;
  const int f0 = (int)tb[0], wb = (int)tb[1], K = wb * 2 * M;
  const float* gw = w + tb[2]; const float* gb = w + tb[3]; const float* Wt = w + tb[4]; const float* bias = w + tb[5];
  float* red = sm + (size_t)K * BS_TT;
  const float* sp = spec + (size_t)b * M * T * F * 2;
  auto feat = [&](int k, int t) -> float {      // k = fl * 2M + ch
    const int fl = k / (2 * M), ch = k - fl * 2 * M, m = ch % M, ri = ch / M;
    return __ldg(&sp[(((size_t)m * T + t) * F + f0 + fl) * 2 + ri]);
  };
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < K * T; i += blockDim.x) { const float v = feat(i % K, i / K); s1 += v; s2 += v * v; }
  const float2 tot = block_sum2(s1, s2, red);
  const float mean = tot.x / (float)(K * T);
  const float rstd = rsqrtf(fmaxf(tot.y / (float)(K * T) - mean * mean, 0.f) + eps);
  const int t0 = blockIdx.y * BS_TT;
  __syncthreads();
  for (int i = threadIdx.x; i < K * BS_TT; i += blockDim.x) {
    const int k = i / BS_TT, tt = i - k * BS_TT;
    sm[i] = t0 + tt < T ? (feat(k, t0 + tt) - mean) * rstd * gw[k] + gb[k] : 0.f;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc[BS_TT];
#pragma unroll
    for (int tt = 0; tt < BS_TT; ++tt) acc[tt] = bias[c];
    for (int k = 0; k < K; ++k) {
      const float wv = __ldg(&Wt[(size_t)k * C + c]);
#pragma unroll
      for (int tt = 0; tt < BS_TT; ++tt) acc[tt] = fmaf(wv, sm[k * BS_TT + tt], acc[tt]);
    }
#pragma unroll
    for (int tt = 0; tt < BS_TT; ++tt)
      if (t0 + tt < T) x[(((size_t)b * T + t0 + tt) * nb + band) * C + c] = acc[tt];
  }
}

// x [B, T, nb, C] -> est [B, S, M, T, F, 2] (= input * mask when masking).  Per band: GroupNorm(1, C) over (C, T),
// Conv1d(C, 4C) -> tanh -> Conv1d(4C, 4C) -> Conv1d(4C, O_b) -> GLU (:221-236, :256-270, :175-182).
__global__ void __launch_bounds__(256) bs_decode_kernel(const float* __restrict__ x, const float* __restrict__ spec, int M,
                                                        int T, int F, int C, int nb, int S,
                                                        const long long* __restrict__ table, const float* __restrict__ w,
                                                        float* __restrict__ est, int masking, float eps) {
  extern __shared__ float sm[];           // xn [C][TT] | h1 [4C][TT] | h2 [4C][TT] | red
  const int band = blockIdx.x, b = blockIdx.z;
  const long long* tb = table + band * 16;
  const int f0 = (int)tb[0], wb = (int)tb[1];
  const int C4 = 4 * C, Q = wb * S * 2 * M;      // Q post-GLU channels, O = 2Q pre-GLU
  const float* gw = w + tb[6]; const float* gb = w + tb[7];
  const float* W1 = w + tb[8]; const float* b1 = w + tb[9];
  const float* W3 = w + tb[10]; const float* b3 = w + tb[11];
  const float* W4 = w + tb[12]; const float* b4 = w + tb[13];
  float* xn = sm; float* h1 = xn + (size_t)C * BS_TT; float* h2 = h1 + (size_t)C4 * BS_TT; float* red = h2 + (size_t)C4 * BS_TT;
  const float* xb = x + (size_t)b * T * nb * C + (size_t)band * C;
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < C * T; i += blockDim.x) {
    const float v = __ldg(&xb[(size_t)(i / C) * nb * C + (i % C)]);
    s1 += v; s2 += v * v;
  }
  const float2 tot = block_sum2(s1, s2, red);
  const float mean = tot.x / (float)(C * T);
  const float rstd = rsqrtf(fmaxf(tot.y / (float)(C * T) - mean * mean, 0.f) + eps);
  const int t0 = blockIdx.y * BS_TT;
  __syncthreads();
  for (int i = threadIdx.x; i < C * BS_TT; i += blockDim.x) {
    const int c = i % C, tt = i / C;
    xn[c * BS_TT + tt] = t0 + tt < T ? (__ldg(&xb[(size_t)(t0 + tt) * nb * C + c]) - mean) * rstd * gw[c] + gb[c] : 0.f;
  }
  __syncthreads();
  auto dense = [&](const float* in, int K, const float* Wt, const float* bias, int N, float* out, bool tanh_act) {
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      float acc[BS_TT];
#pragma unroll
      for (int tt = 0; tt < BS_TT; ++tt) acc[tt] = bias[n];
      for (int k = 0; k < K; ++k) {
        const float wv = __ldg(&Wt[(size_t)k * N + n]);
#pragma unroll
        for (int tt = 0; tt < BS_TT; ++tt) acc[tt] = fmaf(wv, in[k * BS_TT + tt], acc[tt]);
      }
#pragma unroll
      for (int tt = 0; tt < BS_TT; ++tt) out[n * BS_TT + tt] = tanh_act ? tanhf(acc[tt]) : acc[tt];
    }
    __syncthreads();
  };
  dense(xn, C, W1, b1, C4, h1, true);
  dense(h1, C4, W3, b3, C4, h2, false);
  // last layer + GLU + complex mask, one thread per (source, channel, bin): needs the re and im channels and their gates
  const int P = S * M * wb;
  for (int pidx = threadIdx.x; pidx < P; pidx += blockDim.x) {
    const int fl = pidx % wb, m = (pidx / wb) % M, s = pidx / (wb * M);
    const int q_re = ((0 * S + s) * M + m) * wb + fl, q_im = ((1 * S + s) * M + m) * wb + fl;
    float a_re[BS_TT], a_im[BS_TT], g_re[BS_TT], g_im[BS_TT];
#pragma unroll
    for (int tt = 0; tt < BS_TT; ++tt) { a_re[tt] = b4[q_re]; a_im[tt] = b4[q_im]; g_re[tt] = b4[Q + q_re]; g_im[tt] = b4[Q + q_im]; }
    for (int k = 0; k < C4; ++k) {
      const float* wr = W4 + (size_t)k * 2 * Q;
      const float w0 = __ldg(&wr[q_re]), w1 = __ldg(&wr[q_im]), w2 = __ldg(&wr[Q + q_re]), w3 = __ldg(&wr[Q + q_im]);
#pragma unroll
      for (int tt = 0; tt < BS_TT; ++tt) {
        const float hv = h2[k * BS_TT + tt];
        a_re[tt] = fmaf(w0, hv, a_re[tt]); a_im[tt] = fmaf(w1, hv, a_im[tt]);
        g_re[tt] = fmaf(w2, hv, g_re[tt]); g_im[tt] = fmaf(w3, hv, g_im[tt]);
      }
    }
#pragma unroll
    for (int tt = 0; tt < BS_TT; ++tt) {
      const int t = t0 + tt;
      if (t >= T) break;
      float re = a_re[tt] / (1.f + expf(-g_re[tt])), im = a_im[tt] / (1.f + expf(-g_im[tt]));   // GLU(dim=1)
      if (masking) {
        const float2 in = *reinterpret_cast<const float2*>(&spec[((((size_t)b * M + m) * T + t) * F + f0 + fl) * 2]);
        const float r2 = in.x * re - in.y * im, i2 = in.x * im + in.y * re;
        re = r2; im = i2;
      }
      *reinterpret_cast<float2*>(&est[(((((size_t)b * S + s) * M + m) * T + t) * F + f0 + fl) * 2]) = make_float2(re, im);
    }
  }
}

}  // namespace tfl
