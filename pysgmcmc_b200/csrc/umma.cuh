// Thin inline-PTX layer over the Blackwell (sm_100a) 5th-generation tensor core path:
// tcgen05.mma with shared-memory operand descriptors, TMEM accumulators, tcgen05.commit ->
// mbarrier completion, tcgen05.ld for the epilogue.  Only what K14 (csrc/svgd_umma.cu) and
// tools/micro/umma_probe.cu need: one CTA per MMA (cta_group::1), kind::tf32, no swizzle.
//
// Operand layouts ("canonical", no swizzle; a core matrix is 8 rows of 16 bytes):
//   K-major  (A here):  byte(m, k) = (m/8)*SBO + (m%8)*16 + (k/4)*LBO + (k%4)*4
//   MN-major (B here):  byte(n, k) = (n/4)*SBO + (n%4)*4  + (k%8)*16  + (k/8)*LBO
// with SBO / LBO the "stride" / "leading" byte offsets of the 64-bit descriptor
// (bits 0-13 address>>4, 16-29 LBO>>4, 32-45 SBO>>4, 46-47 version = 1, 61-63 swizzle = 0).
// One kind::tf32 instruction contracts K = 8 (two K-major core matrices / one MN-major one).
// The conventions above are verified on the B200 by tools/micro/umma_probe.cu.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sgmcmc {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -----------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- TMEM -----------------------------------------------------------------------------
// whole warp; writes the base address (lane 0, first column) to *slot (shared memory)
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t slot) {
  static_assert(COLS == 32 || COLS == 64 || COLS == 128 || COLS == 256 || COLS == 512, "power of two >= 32");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "n"(COLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void fence_before_thread_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_thread_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// 16 consecutive 32-bit columns of this thread's TMEM lane (warp w of a warpgroup owns lanes
// 32*(w%4) .. +31; taddr = lane << 16 | column)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- descriptors ------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// kind::tf32, fp32 accumulate, dense; majors: 0 = K-major, 1 = MN-major
__host__ __device__ constexpr uint32_t instr_desc_tf32(int m, int n, int a_major, int b_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_major << 15) | ((uint32_t)b_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread for the whole CTA
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// the mbarrier gets one arrival when every tcgen05.mma issued so far by this thread is done
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// fp32 -> (hi, lo) with hi a TF32 value (round to nearest) and lo = x - hi (exact); the tensor
// core truncates lo to TF32 itself.  a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi ("3xTF32").
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  lo = x - hi;
}

}  // namespace umma
}  // namespace sgmcmc
