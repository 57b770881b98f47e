// tf32 tensor-core variants of the fp32 tap-GEMM (kernels_f32.cuh) and of the weight-gradient GEMM (kernels_bwd.cuh) for
// the TRAINING step: mma.sync m16n8k8, tf32 operands (cvt.rna: 2^-11 per operand), fp32 accumulation -- the arithmetic
// the reference trains in at best (configs/musdb18_rtx5090_xlarge.yaml:136 `tf32: true`, under bf16 autocast).  Same
// parameters, same 128 x 128 x 8 block tiles and the same epilogue functors (through their two-column `pair` form: an
// accumulator fragment holds adjacent column pairs).  The fp32 inference / parity mode never runs these kernels.
//   block = 8 warps as 2 (m) x 4 (n); warp tile 64 x 32 = 4 x 4 fragments, 16 MMAs per 8-deep k step and 24 shared loads.
//   shared tiles are k-major with a pitch of 136 floats: fragment loads (k = t or t + 4, m / n = g) hit 32 distinct banks.
#pragma once
#include "kernels_bwd.cuh"

namespace tfl {

constexpr int MMA_PITCH = GBM + 8;
constexpr int MMA_BK = 16;   // k depth per barrier round (two m16n8k8 steps): the 8-deep version ran latency-bound (r02: 89 TFLOP/s)

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// one 8-deep k step of the warp's 64 x 32 tile: As / Bs are [8][MMA_PITCH] tiles holding tf32 bit patterns
__device__ __forceinline__ void mma_warp_step(const uint32_t* As, const uint32_t* Bs, float (&acc)[4][4][4], int wm, int wn, int g, int t) {
  uint32_t a[4][4], b[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = wm * 64 + i * 16 + g;
    a[i][0] = As[t * MMA_PITCH + m]; a[i][1] = As[t * MMA_PITCH + m + 8];
    a[i][2] = As[(t + 4) * MMA_PITCH + m]; a[i][3] = As[(t + 4) * MMA_PITCH + m + 8];
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = wn * 32 + j * 8 + g;
    b[j][0] = Bs[t * MMA_PITCH + n]; b[j][1] = Bs[(t + 4) * MMA_PITCH + n];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) mma_tf32(acc[i][j], a[i], b[j]);
}

template <class Epi>
__global__ void __launch_bounds__(256) tap_gemm_mma_kernel(TapGemm p, Epi epi) {
  __shared__ __align__(16) uint32_t As[2][MMA_BK][MMA_PITCH];
  __shared__ __align__(16) uint32_t Bs[2][MMA_BK][MMA_PITCH];
  __shared__ long long row_base[GBM];
  __shared__ int row_j[GBM], row_s[GBM];
  const int tid = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * GBM;
  const int n0 = blockIdx.y * GBN;
  if (tid < GBM) {
    const long long r = m0 + tid;
    if (r < p.M) {
      const int s = (int)(r / p.Sout), j = (int)(r - (long long)s * p.Sout);
      row_s[tid] = s; row_j[tid] = j; row_base[tid] = p.amap.base(s);
    } else { row_s[tid] = -1; row_j[tid] = 0; row_base[tid] = 0; }
  }
  __syncthreads();
  const int a_row = tid >> 1, a_kq = (tid & 1) << 2;       // A: row a_row, k = a_kq .. +3 and a_kq + 8 .. +11
  const int b_row = tid >> 5, b_col = (tid & 31) << 2;    // B: k rows b_row and b_row + 8, columns b_col .. +3
  const int kt_per_tap = p.Kc / MMA_BK, n_kt = p.taps * kt_per_tap;
  const int my_s = row_s[a_row], my_j = row_j[a_row];
  const long long my_base = row_base[a_row];
  auto load_a = [&](int kt, float4 (&a)[2]) {
    const int tap = kt / kt_per_tap, c0 = (kt - tap * kt_per_tap) * MMA_BK;
    const int pos = my_j + tap - p.padL;
    if (my_s < 0 || pos < 0 || pos >= p.Sin) { a[0] = a[1] = make_float4(0.f, 0.f, 0.f, 0.f); return; }
    const float* src = p.A + my_base + (long long)pos * p.amap.pos_stride + c0 + a_kq;
    a[0] = __ldg(reinterpret_cast<const float4*>(src));
    a[1] = __ldg(reinterpret_cast<const float4*>(src + 8));
  };
  auto load_b = [&](int kt, float4 (&b)[2]) {
    const int n = n0 + b_col;
    if (n >= p.N) { b[0] = b[1] = make_float4(0.f, 0.f, 0.f, 0.f); return; }
    const float* src = p.W + ((size_t)kt * MMA_BK + b_row) * p.N + n;
    b[0] = __ldg(reinterpret_cast<const float4*>(src));
    b[1] = __ldg(reinterpret_cast<const float4*>(src + (size_t)8 * p.N));
  };
  auto stash = [&](int buf, const float4 (&a)[2], const float4 (&b)[2]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      As[buf][8 * h + a_kq + 0][a_row] = to_tf32(a[h].x); As[buf][8 * h + a_kq + 1][a_row] = to_tf32(a[h].y);
      As[buf][8 * h + a_kq + 2][a_row] = to_tf32(a[h].z); As[buf][8 * h + a_kq + 3][a_row] = to_tf32(a[h].w);
      *reinterpret_cast<uint4*>(&Bs[buf][8 * h + b_row][b_col]) = make_uint4(to_tf32(b[h].x), to_tf32(b[h].y), to_tf32(b[h].z), to_tf32(b[h].w));
    }
  };
  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, t = lane & 3;
  float4 na[2], nb[2];
  load_a(0, na); load_b(0, nb);
  stash(0, na, nb);
  __syncthreads();
  for (int kt = 0; kt < n_kt; ++kt) {
    const int cur = kt & 1;
    const bool more = kt + 1 < n_kt;
    if (more) { load_a(kt + 1, na); load_b(kt + 1, nb); }
    mma_warp_step(&As[cur][0][0], &Bs[cur][0][0], acc, wm, wn, g, t);
    mma_warp_step(&As[cur][8][0], &Bs[cur][8][0], acc, wm, wn, g, t);
    if (more) stash(cur ^ 1, na, nb);
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + wn * 32 + j * 8 + 2 * t;
    if (n >= p.N) continue;
    const float b0 = p.bias != nullptr ? __ldg(&p.bias[n]) : 0.f, b1 = p.bias != nullptr ? __ldg(&p.bias[n + 1]) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int rl = wm * 64 + i * 16 + g + 8 * h;
        const int s = row_s[rl];
        if (s < 0) continue;
        epi.pair(s, row_j[rl], m0 + rl, n, p.N, acc[i][j][2 * h] + b0, acc[i][j][2 * h + 1] + b1);
      }
    }
  }
}

__global__ void __launch_bounds__(256) tap_wgrad_mma_kernel(TapWgrad p) {
  __shared__ __align__(16) uint32_t As[2][MMA_BK][MMA_PITCH];
  __shared__ __align__(16) uint32_t Bs[2][MMA_BK][MMA_PITCH];
  const int tid = threadIdx.x;
  const int tiles_i = (p.Kc + GBM - 1) / GBM, tiles_n = (p.N + GBN - 1) / GBN;
  int bx = blockIdx.x;
  const int tn = bx % tiles_n; bx /= tiles_n;
  const int ti = bx % tiles_i; bx /= tiles_i;
  const int tap = bx;
  const int i0 = ti * GBM, n0 = tn * GBN;
  const long long r_lo = (long long)blockIdx.y * p.rows_per_split;
  const long long r_hi = r_lo + p.rows_per_split < p.R ? r_lo + p.rows_per_split : p.R;
  if (r_lo >= r_hi) return;
  const int l_row = tid >> 5, l_col = (tid & 31) << 2;
  auto load = [&](long long r0, float4 (&a)[2], float4 (&b)[2]) {   // rows r0 + l_row and r0 + l_row + 8
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      a[h] = make_float4(0.f, 0.f, 0.f, 0.f); b[h] = a[h];
      const long long r = r0 + l_row + 8 * h;
      if (r >= r_hi) continue;
      const int s = (int)(r / p.Sout), j = (int)(r - (long long)s * p.Sout);
      const int pos = j + tap - p.padL;
      if (pos >= 0 && pos < p.Sin && i0 + l_col < p.Kc)
        a[h] = __ldg(reinterpret_cast<const float4*>(p.A + p.amap.base(s) + (long long)pos * p.amap.pos_stride + i0 + l_col));
      if (n0 + l_col < p.N)
        b[h] = __ldg(reinterpret_cast<const float4*>(p.B + p.bmap.base(s) + (long long)j * p.bmap.pos_stride + n0 + l_col));
    }
  };
  auto stash = [&](int buf, const float4 (&a)[2], const float4 (&b)[2]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      *reinterpret_cast<uint4*>(&As[buf][l_row + 8 * h][l_col]) = make_uint4(to_tf32(a[h].x), to_tf32(a[h].y), to_tf32(a[h].z), to_tf32(a[h].w));
      *reinterpret_cast<uint4*>(&Bs[buf][l_row + 8 * h][l_col]) = make_uint4(to_tf32(b[h].x), to_tf32(b[h].y), to_tf32(b[h].z), to_tf32(b[h].w));
    }
  };
  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, t = lane & 3;
  float4 a[2], b[2];
  load(r_lo, a, b);
  stash(0, a, b);
  __syncthreads();
  int cur = 0;
  for (long long r0 = r_lo; r0 < r_hi; r0 += MMA_BK) {
    const bool more = r0 + MMA_BK < r_hi;
    if (more) load(r0 + MMA_BK, a, b);
    mma_warp_step(&As[cur][0][0], &Bs[cur][0][0], acc, wm, wn, g, t);
    mma_warp_step(&As[cur][8][0], &Bs[cur][8][0], acc, wm, wn, g, t);
    if (more) stash(cur ^ 1, a, b);
    __syncthreads();
    cur ^= 1;
  }
  float* o = p.out + (size_t)tap * p.Kc * p.N;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int ii = i0 + wm * 64 + i * 16 + g + 8 * h;
      if (ii >= p.Kc) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int nn = n0 + wn * 32 + j * 8 + 2 * t;
        if (nn < p.N) {
          atomicAdd(&o[(size_t)ii * p.N + nn], acc[i][j][2 * h]);
          atomicAdd(&o[(size_t)ii * p.N + nn + 1], acc[i][j][2 * h + 1]);
        }
      }
    }
}

}  // namespace tfl
