// sm_100a building blocks written as inline PTX: mbarrier, 1-D bulk async copy (TMA engine,
// UBLKCP in SASS), tcgen05 alloc / mma / commit / ld, UMMA shared-memory and instruction
// descriptors.  No CUTLASS: the bit layouts follow the PTX ISA "tcgen05 matrix descriptor" /
// "instruction descriptor" tables.
//
// Shared-memory operand layout used everywhere ("chunk-major", the canonical no-swizzle
// K-major layout): an operand tile with R rows and Kd bf16 columns is stored as
//     byte_offset(r, c) = ((c / 8) * R + r) * 16 + (c % 8) * 2
// i.e. 8x8 core matrices of 128 contiguous bytes, 16 B per row, consecutive rows adjacent.
//   * as a K-major operand:   LBO = R * 16 (next 8 columns), SBO = 128 (next 8 rows)
//   * as an MN-major operand (rows are the K index, columns the N index; V in P.V):
//                             LBO = 128 (next 8 rows of K), SBO = R * 16 (next 8 columns of N)
// Because rows sit at a uniform 16-byte pitch, a descriptor whose start address is advanced by
// t * 16 bytes addresses the same tile shifted down by t rows -- which is how the K taps of the
// conv / transposed conv are fed to the tensor core from ONE staged tile (implicit GEMM without
// im2col copies).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace tfl {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a fully converged warp.  Single-thread instructions (tcgen05.mma / commit, bulk copies) are
// issued under this predicate while the WHOLE warp runs the surrounding control flow: ptxas then keeps
// descriptors in uniform registers instead of wrapping every UTCHMMA in an elect/retry loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must never hang the GPU, and must never pass silently either.  The first wait that
// expires records (block, thread, barrier address, parity) in g_wait_timeout AND in a host-mapped copy of that record
// (g_timeout_host, installed per device by the library; it survives the trap), then traps: the launch fails, the
// context reports cudaErrorLaunchFailed at the next synchronising call, and every later tfl_* entry point returns an
// error naming the record (tfl_api.cu: timeout_pending).
__device__ unsigned int g_wait_timeout[5] = {0, 0, 0, 0, 0};
__device__ unsigned int* g_timeout_host = nullptr;
constexpr long long WAIT_LIMIT_CLOCKS = 4000000000LL;         // ~2 s at 1.9 GHz; legitimate waits are < 1 ms
__device__ __forceinline__ void wait_expired(uint32_t bar, uint32_t parity) {
  if (atomicCAS(&g_wait_timeout[0], 0u, 1u) == 0u) {
    g_wait_timeout[1] = blockIdx.x; g_wait_timeout[2] = threadIdx.x; g_wait_timeout[3] = bar; g_wait_timeout[4] = parity;
    unsigned int* h = g_timeout_host;
    if (h != nullptr) {
      h[1] = blockIdx.x; h[2] = threadIdx.x; h[3] = bar; h[4] = parity;
      __threadfence_system();
      *(volatile unsigned int*)h = 1u;
    }
    __threadfence_system();
  }
  __trap();
}
__device__ __forceinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 64; ++i)
      if (mbar_try_wait(bar, parity)) return;
    // (no global memory in the polling loop: a ~1.5 k clk load per 64 polls sat between every producer -> consumer
    // handoff of the pipelines)
    if (clock64() - t0 > WAIT_LIMIT_CLOCKS) wait_expired(bar, parity);
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // fast path: a handful of polls inline (try_wait itself suspends the thread for a HW-defined time slice)
  if (mbar_try_wait(bar, parity)) return;
  if (mbar_try_wait(bar, parity)) return;
  if (mbar_try_wait(bar, parity)) return;
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity);
}

// Optional event trace (diagnostic): when tfl_debug_set_trace() has installed a buffer, block 0 of the traced kernels
// stores clock64() stamps at [event * 64 + index].
__device__ unsigned long long* g_trace = nullptr;
__device__ int g_trace_base = 0;       // first chunk / tile index recorded (tfl_debug_set_option(TFL_OPT_TRACE_BASE, n))
__device__ __forceinline__ void trace_event(unsigned long long* tr, int event, int index) {
  if (tr == nullptr) return;
  index -= g_trace_base;
  if (index >= 0 && index < 64) tr[event * 64 + index] = (unsigned long long)clock64();
}

__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy st.shared -> visible to the async proxy (tcgen05.mma / bulk copy reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- 1-D bulk async copy global -> shared, completion on an mbarrier ----------------------
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// bulk prefetch of a contiguous global range into L2 (no registers, no shared memory, no completion to wait for):
// used to pull the fp32 residual rows a producer warp will read two tiles from now, so that its loads -- whose number
// in flight is bounded by registers -- hit L2 instead of HBM.  16-byte aligned address, size a multiple of 16.
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

// ---- tcgen05 ------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; one thread issues.
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// mbarrier arrives once all tcgen05.mma issued so far by this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// UMMA shared-memory matrix descriptor, SWIZZLE_NONE ("interleave"), descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, M x N tile; b_mn_major = B stored N-contiguous.
__host__ __device__ constexpr uint32_t instr_desc(int m, int n, bool b_mn_major = false, bool a_mn_major = false) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// registers -> TMEM (used to stage softmax probabilities as the A operand of P.V)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace tc
}  // namespace tfl
