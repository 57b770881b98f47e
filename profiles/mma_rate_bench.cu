// Micro-benchmark: tcgen05.mma issue rate as a function of the A-operand start alignment (row-shifted descriptors
// of the implicit-GEMM conv taps), the N extent and the accumulate chain.  One CTA per SM, one issuing warp.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../mss_tf_locoformer_b200/csrc -o mma_rate_bench mma_rate_bench.cu
//
// Prints SM clocks per MMA (M128 x N x K16, bf16; the tensor-pipe floor is N/2 clocks).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "kernels_tc.cuh"

using namespace tfl::tc;
using tfl::mma_burst;

// shift_rows: A descriptor start advanced by shift_rows * 16 B; taps: cycle through shifts 0..taps-1 (like the conv)
__global__ void __launch_bounds__(512, 1) rate_kernel(int N, int n_mma, int shift_rows, int taps, int b_bytes_step,
                                                      int extra_sts, int ld_tmem, const char* img, int stream_stages, int pattern, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int AR = AR_ROWS;                                 // rows per A tile (128 + up to 8 shift rows)
  const uint32_t a_bytes = 16u * AR * 16;              // 128 bf16 columns, chunk-major
  const uint32_t sa = smem_u32(smem), sb = sa + ((a_bytes + 1023) & ~1023u);
  const uint32_t bar_ = sb + 4 * 256 * 32 * 2 + 1024;   // after 64 KB of B
  const uint32_t bar = bar_;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (bar - sa) + 128);
  const int warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < (bar - sa) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { for (int w = 0; w < 2; ++w) { mbar_init(bar + 32 * w, 1); mbar_init(bar + 32 * w + 16, 1 << 20); mbar_init(bar + 32 * w + 24, 1); } fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0 || (warp == 1 && pattern == 4)) {
    const uint32_t idesc = instr_desc(128, N);
    long long t0 = 0;
    const uint32_t bar = bar_ + 32 * warp;
    for (int rep = 0; rep < 2; ++rep) {                // rep 0 warms up
      __syncwarp();
      t0 = clock64();
      const uint32_t hi = (128u >> 4) | (1u << 14);
      const uint32_t lo_a = (uint32_t)AR << 16, lo_b = (uint32_t)N << 16;
      const uint32_t a16 = (sa >> 4) + shift_rows, b16 = sb >> 4;
      const uint32_t bar2 = bar + 16, bar3 = bar + 24;
      if (pattern == 0) {
        if (elect_one()) {
          for (int o = 0; o < n_mma / 8; ++o) {              // 8 K steps of one tap per iteration
            const uint32_t tap = o & (taps - 1);
            const uint32_t bo = (o & 7) * (b_bytes_step >> 4);
            const uint32_t d = tmem + (o & 1) * 256;
            mma_burst<4>(d, (a16 + tap) | lo_a, (b16 + bo) | lo_b, hi, idesc, o >= 2, 2u * AR, 2u * N);
            mma_burst<4>(d, (a16 + tap + 8u * AR) | lo_a, (b16 + bo + 8u * N) | lo_b, hi, idesc, 1u, 2u * AR, 2u * N);
          }
          mma_commit(bar);
        }
      } else {
        // the issue pattern of ffn_tc_kernel: per weight stage, 4 MMAs for tile 0 then 4 for tile 1 (same B), each under
        // its own elect, then a commit; optionally a wait on an (always complete) barrier + fence per stage
        if (rep == 0 && elect_one()) { mbar_arrive(bar3); }
        __syncwarp();
        for (int o = 0; o < n_mma / 8; ++o) {
          if (pattern >= 3) { mbar_wait(bar3, 0); tc_fence_after(); }
          const uint32_t tap = (o >> 1) & (taps - 1);
          const uint32_t bo = (o & 7) * (b_bytes_step >> 4);
          for (int t = 0; t < (pattern == 4 ? 1 : 2); ++t) {
            if (elect_one())
              tfl::mma_run(n_mma >= 0 ? 4 : 3, tmem + t * 128 + warp * 256, (a16 + tap + (o & 1) * 8u * AR) | lo_a, (b16 + bo) | lo_b, hi, idesc, o >= 2, 2u * AR, 2u * N);
            __syncwarp();
          }
          if (pattern >= 2) { if (elect_one()) mma_commit(bar2); __syncwarp(); }
        }
        if (elect_one()) mma_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, rep & 1);
      tc_fence_after();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
  } else if (warp == 6 && stream_stages > 0) {
    // concurrent weight streaming: bulk copies of 16 KB stages into a 3-slot ring at a paced rate (no consumer)
    const uint32_t ring = bar + 1024 + 96 * 8 * 16, rb = bar + 256;
    if (threadIdx.x == 192) {
      for (int i = 0; i < 3; ++i) mbar_init(rb + 8 * i, 1);
      fence_barrier_init();
      for (int s = 0; s < stream_stages; ++s) {
        const int slot = s % 3;
        if (s >= 3) mbar_wait(rb + 8 * slot, ((s / 3) - 1) & 1);
        mbar_arrive_expect_tx(rb + 8 * slot, 16384);
        bulk_g2s(ring + slot * 16384, img + (size_t)(s % 64) * 16384, 16384, rb + 8 * slot);
      }
      for (int s = stream_stages; s < stream_stages + 3; ++s) if (s >= 3) mbar_wait(rb + 8 * (s % 3), ((s / 3) - 1) & 1);
    }
  } else if (warp >= 2 && extra_sts) {
    // other warps hammer shared memory with stores (the epilogue / producer traffic of the real kernel)
    const uint32_t p = bar + 1024;
    for (int i = 0; i < extra_sts; ++i)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(p + 16u * (((threadIdx.x - 64) & 127) + 64 * (i & 7))), "r"(i) : "memory");
  } else if (warp >= 2 && ld_tmem < 0) {
    // busy ALU / MUFU neighbours on every scheduler (the SwiGLU epilogue warps of the real kernel)
    float a = threadIdx.x * 0.001f, b = 1.0001f;
    for (int i = 0; i < -ld_tmem; ++i) {
#pragma unroll
      for (int u = 0; u < 16; ++u) { a = fmaf(a, b, 0.5f); b = fmaf(b, 0.999f, a * 1e-9f); }
      float t; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(a)); a += t;
    }
    if (a == 123.f) out[1] = 1;
  } else if (warp >= 2 && ld_tmem) {
    // TMEM reads of the accumulator columns while the MMAs run (the epilogue's tcgen05.ld traffic)
    uint32_t r[32];
    uint32_t acc = 0;
    for (int i = 0; i < ld_tmem; ++i) {
      tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 32 * (i & 7), r);
      tc_wait_ld();
      acc += r[i & 31];
    }
    if (acc == 0x12345678u) out[0] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* out; cudaMalloc(&out, 148 * sizeof(long long));
  long long h[148];
  const int n_mma = 2048;
  const size_t smem = 16 * AR_ROWS * 16 + 1024 + 65536 + 2048 + 96 * 8 * 16 + 1024 + 3 * 16384;
  char* img; cudaMalloc(&img, 64 * 16384); cudaMemset(img, 0, 64 * 16384);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  printf("%5s %6s %5s %7s %6s %6s %6s %10s\n", "N", "shift", "taps", "b_step", "sts", "ldtm", "stream/pattern", "clk/MMA");
  struct Cfg { int N, shift, taps, bstep, sts, ldt, stream, pat; };
  const Cfg cfgs[] = {
      {128, 0, 4, 4096, 0, 0, 0, 0}, {64, 0, 4, 4096, 0, 0, 0, 0}, {32, 0, 4, 4096, 0, 0, 0, 0}, {16, 0, 4, 4096, 0, 0, 0, 0},
      {128, 0, 4, 4096, 0, 0, 0, 3}, {64, 0, 4, 4096, 0, 0, 0, 3}, {32, 0, 4, 4096, 0, 0, 0, 3},
      {128, 0, 4, 4096, 0, 0, 0, 4}, {64, 0, 4, 4096, 0, 0, 0, 4}, {32, 0, 4, 4096, 0, 0, 0, 4},
      {128, 0, 4, 4096, 0, -20000, 0, 0}, {128, 0, 4, 4096, 0, -20000, 0, 3}, {128, 0, 4, 4096, 0, -20000, 0, 4},
      {32, 0, 4, 4096, 0, -20000, 0, 3}, {32, 0, 4, 4096, 0, -20000, 0, 4},
  };
  for (const Cfg& c : cfgs) {
    rate_kernel<<<148, 512, smem>>>(c.N, n_mma, c.shift, c.taps, c.bstep, c.sts, c.ldt, img, c.stream, c.pat, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%5d %6d %5d %7d %6d %6d %6d %10.1f\n", c.N, c.shift, c.taps, c.bstep, c.sts, c.ldt, c.stream * 10 + c.pat, (double)mx / n_mma);
  }
  return 0;
}
