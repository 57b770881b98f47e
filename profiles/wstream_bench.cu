// Micro-benchmark behind the FFN weight-ring design: how fast can every SM stream the same ~1.2 MB weight image
// out of L2 through a ring of NS bulk-copy stages, as a function of ring depth, consumer time per stage (the MMA
// time the stage feeds) and 2-CTA cluster multicast?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o wstream_bench wstream_bench.cu && ./wstream_bench
//
// Prints SM clocks per stage for each configuration (one loader lane, one consumer lane per CTA).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(rank) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CL>
__global__ void __launch_bounds__(64, 1) stream_kernel(const char* img, uint32_t image_bytes, uint32_t stage_bytes, int NS,
                                                       int n_pass, int delay, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + NS * stage_bytes;
  auto FULL = [&](int i) { return bar0 + 8u * i; };
  auto EMPTY = [&](int i) { return bar0 + 8u * (16 + i); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CL > 1 ? cluster_rank() : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(FULL(i), 1); mbar_init(EMPTY(i), CL); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (CL > 1) cluster_sync();
  const int stages = image_bytes / stage_bytes;
  const long long t0 = clock64();
  if (warp == 0 && lane == 0) {
    uint32_t slot = 0, ph = 0;
    for (int pass = 0; pass < n_pass; ++pass)
      for (int s = 0; s < stages; ++s) {
        mbar_wait(EMPTY(slot), ph ^ 1);
        mbar_expect_tx(FULL(slot), stage_bytes);
        const uint32_t dst = sbase + slot * stage_bytes;
        const char* src = img + (size_t)s * stage_bytes;
        if (CL == 1) bulk_g2s(dst, src, stage_bytes, FULL(slot));
        else {
          const uint32_t part = stage_bytes / CL;
          bulk_g2s_mc(dst + rank * part, src + rank * part, part, FULL(slot), (uint16_t)((1u << CL) - 1));
        }
        if (++slot == (uint32_t)NS) { slot = 0; ph ^= 1; }
      }
  } else if (warp == 1 && lane == 0) {
    uint32_t slot = 0, ph = 0;
    long long t_done = clock64();
    for (int pass = 0; pass < n_pass; ++pass)
      for (int s = 0; s < stages; ++s) {
        mbar_wait(FULL(slot), ph);
        const long long now = clock64();
        t_done = (now > t_done ? now : t_done) + delay;     // in-order consumer: `delay` clocks of MMA time per stage
        while (clock64() < t_done) {}
        if (CL == 1) mbar_arrive(EMPTY(slot));
        else for (uint32_t r = 0; r < CL; ++r) mbar_arrive_remote(EMPTY(slot), r);
        if (++slot == (uint32_t)NS) { slot = 0; ph ^= 1; }
      }
    out[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (CL > 1) cluster_sync();
}

int main() {
  const uint32_t image = 1179648;  // 6 chunks x (8 + 2) stages x 16 KB + ... ~ the Variant D FFN image (72 x 16 KB)
  char* img; long long* out;
  cudaMalloc(&img, image); cudaMemset(img, 1, image);
  cudaMalloc(&out, 148 * sizeof(long long));
  long long h[148];
  const int n_pass = 20;
  printf("%8s %6s %4s %6s %10s %10s\n", "cluster", "stage", "NS", "delay", "clk/stage", "B/clk/SM");
  for (int cl = 1; cl <= 2; ++cl)
    for (uint32_t stage : {16384u, 8192u})
      for (int NS : {3, 5, 7, 10, 14})
        for (int delay : {0, 512}) {
          if (NS * stage > 200 * 1024) continue;
          const size_t smem = (size_t)NS * stage + 512;
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(148); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
          cfg.attrs = at; cfg.numAttrs = 1;
          cudaError_t e;
          const int d = delay * (int)(stage / 1024) / 16;   // MMA time scales with the stage size
          if (cl == 1) {
            cudaFuncSetAttribute(stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            e = cudaLaunchKernelEx(&cfg, stream_kernel<1>, (const char*)img, image, stage, NS, n_pass, d, out);
          } else {
            cudaFuncSetAttribute(stream_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            e = cudaLaunchKernelEx(&cfg, stream_kernel<2>, (const char*)img, image, stage, NS, n_pass, d, out);
          }
          if (e != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
          long long mx = 0;
          for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
          const double per_stage = (double)mx / ((double)n_pass * (image / stage));
          printf("%8d %6u %4d %6d %10.1f %10.2f\n", cl, stage, NS, d, per_stage, stage / per_stage);
        }
  return 0;
}
