// Stand-alone check of the cta_group::2 plumbing planned for the round-2 FFN kernel: a 2-CTA cluster computes
// D[256, N] = A[256, Kd] . B[N, Kd]^T with ONE tcgen05.mma.cta_group::2 chain issued by the leader CTA.
//   * each CTA stages its own 128 rows of A and its own HALF of B's N rows (chunk-major, no swizzle),
//   * TMEM is allocated with tcgen05.alloc.cta_group::2 by the same warp of both CTAs,
//   * completion is signalled to both CTAs with tcgen05.commit.cta_group::2 ... .multicast::cluster,
//   * each CTA reads back its 128 rows of D.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../mss_tf_locoformer_b200/csrc -I../include -o cta2_selftest cta2_selftest.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "tc_common.cuh"

using namespace tfl::tc;

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mma_ss2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit2(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
cta2_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int Kd, long long* rate_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t rank = cluster_rank();
  const int NH = N / 2;
  const uint32_t a_bytes = (uint32_t)(Kd / 8) * 128 * 16, b_bytes = (uint32_t)(Kd / 8) * NH * 16;
  uint8_t* sa = smem;
  uint8_t* sb = smem + a_bytes;
  const uint32_t bar = smem_u32(sb + b_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sb + b_bytes + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 128 * Kd; i += blockDim.x) {
    const int r = i / Kd, c = i % Kd;
    *reinterpret_cast<__nv_bfloat16*>(sa + ((size_t)(c >> 3) * 128 + r) * 16 + (c & 7) * 2) = __float2bfloat16_rn(A[(size_t)(rank * 128 + r) * Kd + c]);
  }
  for (int i = threadIdx.x; i < NH * Kd; i += blockDim.x) {
    const int r = i / Kd, c = i % Kd;
    *reinterpret_cast<__nv_bfloat16*>(sb + ((size_t)(c >> 3) * NH + r) * 16 + (c & 7) * 2) = __float2bfloat16_rn(B[(size_t)(rank * NH + r) * Kd + c]);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (rank == 0 && warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = instr_desc(256, N);
      for (int kk = 0; kk < Kd / 16; ++kk) {
        const uint64_t ad = smem_desc(smem_u32(sa) + kk * 2 * 128 * 16, 128 * 16, 128);
        const uint64_t bd = smem_desc(smem_u32(sb) + kk * 2 * NH * 16, NH * 16, 128);
        mma_ss2(tmem, ad, bd, idesc, kk != 0);
      }
      mma_commit2(bar, 3);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0);
  if (rate_out != nullptr) {                 // issue-rate measurement: 1024 more MMAs into columns 256.., timed
    tc_fence_after();
    long long t0 = clock64();
    if (rank == 0 && warp == 0) {
      if (elect_one()) {
        const uint32_t idesc = instr_desc(256, N);
        for (int i = 0; i < 1024; ++i) {
          const int kk = i & 7;
          const uint64_t ad = smem_desc(smem_u32(sa) + kk * 2 * 128 * 16, 128 * 16, 128);
          const uint64_t bd = smem_desc(smem_u32(sb) + kk * 2 * NH * 16, NH * 16, 128);
          mma_ss2(tmem + 256, ad, bd, idesc, i != 0);
        }
        mma_commit2(bar, 3);
      }
      __syncwarp();
    }
    mbar_wait(bar, 1);
    if (threadIdx.x == 0) rate_out[rank] = clock64() - t0;
  }
  tc_fence_after();
  const int m = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tc_wait_ld();
    for (int e = 0; e < 16; ++e) D[(size_t)(rank * 128 + m) * N + c0 + e] = __uint_as_float(r[e]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

int main() {
  const int N = 128, Kd = 128;
  std::vector<float> A(256 * Kd), B(N * Kd), D(256 * N), W(256 * N);
  srand(1);
  for (auto& v : A) v = (rand() % 2001 - 1000) / 1000.f;
  for (auto& v : B) v = (rand() % 2001 - 1000) / 1000.f;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < Kd; ++k) s += (double)bf(A[m * Kd + k]) * bf(B[n * Kd + k]);
      W[m * N + n] = (float)s;
    }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, D.size() * 4);
  const size_t smem = (size_t)(Kd / 8) * 128 * 16 + (size_t)(Kd / 8) * (N / 2) * 16 + 64;
  cudaFuncSetAttribute(cta2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long* dR; cudaMalloc(&dR, 16);
  cta2_kernel<<<2, 128, smem>>>(dA, dB, dD, N, Kd, dR);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double worst[2] = {0, 0};
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) worst[m / 128] = fmax(worst[m / 128], fabs((double)D[m * N + n] - W[m * N + n]));
  unsigned int to[5]; cudaMemcpyFromSymbol(to, tfl::tc::g_wait_timeout, sizeof(to));
  printf("cta_group::2 M256 x N%d x K%d: max |err| rows 0-127 %.3e, rows 128-255 %.3e (timeout flag %u)\n", N, Kd, worst[0], worst[1], to[0]);
  long long hr[2]; cudaMemcpy(hr, dR, 16, cudaMemcpyDeviceToHost);
  printf("issue rate: %.1f clk per tcgen05.mma.cta_group::2 (M256 x N%d x K16), leader %lld peer %lld clk for 1024\n", hr[0] / 1024.0, N, hr[0], hr[1]);
  printf("sample D[0][0]=%.4f want %.4f; D[200][77]=%.4f want %.4f\n", D[0], W[0], D[200 * N + 77], W[200 * N + 77]);
  return worst[0] < 1e-2 && worst[1] < 1e-2 ? 0 : 2;
}
