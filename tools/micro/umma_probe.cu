// Probe: pins the tcgen05.mma (kind::tf32, cta_group::1, no swizzle) operand-descriptor
// conventions that csrc/umma.cuh states, on the real B200, with exact small-integer matrices.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/micro/_bin/umma_probe tools/micro/umma_probe.cu
//   tools/micro/_bin/umma_probe <case>      (one case per process: a faulting case cannot poison the next)
// case bit 0: swap the LBO/SBO fields of A's descriptor; bit 1: swap them for B;
// bit 2: B is K-major (same layout rule as A) instead of MN-major; bit 3 (with bit 2): B's
// 8-row groups are padded to SBO = 144 bytes (K14's bank-conflict-free staging layout).
// Result on B200 (profiles/r01_umma_probe.jsonl): case 4 and case 12 are exact; swapped
// fields fault; the MN-major attempt returns zeros (that layout is NOT what this file assumes).
// Each case runs D = A[128 x 16] * B[256 x 16]^T as two K = 8 instructions (the second one
// accumulates and uses descriptors advanced by one k-step) and compares with the host product.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../pysgmcmc_b200/csrc/umma.cuh"

using namespace sgmcmc::umma;

constexpr int M = 128, N = 256, K = 16;
constexpr uint32_t A_SBO = 128, A_LBO = 2048;          // K-major, 128 rows
constexpr uint32_t BMN_SBO = 128, BMN_LBO = 8192;      // MN-major, 256 columns: 64 groups of 4
constexpr uint32_t BK_SBO = 128, BK_LBO = 4096;        // K-major, 256 rows
constexpr uint32_t BP_SBO = 144, BP_LBO = 32 * 144;    // K-major, 256 rows, padded row groups
constexpr uint32_t A_BYTES = A_LBO * (K / 4), B_BYTES = BP_LBO * (K / 4);

__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, int swap_a, int swap_b,
                                             int b_kmajor, int b_padded) {
  const uint32_t BK_SBO = b_padded ? BP_SBO : ::BK_SBO, BK_LBO = b_padded ? BP_LBO : ::BK_LBO;
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  float* sa = reinterpret_cast<float*>(smem);
  float* sb = reinterpret_cast<float*>(smem + A_BYTES);
  const int tid = threadIdx.x;
  for (int i = tid; i < (int)((A_BYTES + B_BYTES) / 4); i += 128) reinterpret_cast<float*>(smem)[i] = 0.0f;
  __syncthreads();
  for (int i = tid; i < M * K; i += 128) {
    const int m = i / K, k = i % K;
    sa[((m / 8) * A_SBO + (m % 8) * 16 + (k / 4) * A_LBO + (k % 4) * 4) / 4] = A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    const uint32_t off = b_kmajor ? (n / 8) * BK_SBO + (n % 8) * 16 + (k / 4) * BK_LBO + (k % 4) * 4
                                  : (n / 4) * BMN_SBO + (n % 4) * 4 + (k % 8) * 16 + (k / 8) * BMN_LBO;
    sb[off / 4] = B[i];
  }
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init_fence();
  }
  if (tid < 32) tmem_alloc<256>(smem_u32(&tmem_slot));
  fence_before_thread_sync();
  __syncthreads();
  fence_after_thread_sync();
  const uint32_t taddr = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = instr_desc_tf32(M, N, 0, b_kmajor ? 0 : 1);
    for (int ks = 0; ks < 2; ++ks) {
      const uint32_t a_addr = smem_u32(sa) + ks * 2 * A_LBO;
      const uint32_t b_addr = smem_u32(sb) + (b_kmajor ? ks * 2 * BK_LBO : ks * BMN_LBO);
      const uint32_t b_lbo = b_kmajor ? BK_LBO : BMN_LBO, b_sbo = b_kmajor ? BK_SBO : BMN_SBO;
      const uint64_t da = swap_a ? smem_desc(a_addr, A_SBO, A_LBO) : smem_desc(a_addr, A_LBO, A_SBO);
      const uint64_t db = swap_b ? smem_desc(b_addr, b_sbo, b_lbo) : smem_desc(b_addr, b_lbo, b_sbo);
      mma_tf32(taddr, da, db, idesc, ks > 0);
    }
    commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  fence_after_thread_sync();
  const int warp = tid / 32, lane = tid % 32;
  for (int c = 0; c < N / 16; ++c) {
    float v[16];
    tmem_ld16(taddr + ((uint32_t)(warp * 32) << 16) + c * 16, v);
    for (int e = 0; e < 16; ++e) D[(warp * 32 + lane) * N + c * 16 + e] = v[e];
  }
  fence_before_thread_sync();
  __syncthreads();
  if (tid < 32) {
    fence_after_thread_sync();
    tmem_dealloc<256>(taddr);
  }
}

int main(int argc, char** argv) {
  const int which = argc > 1 ? atoi(argv[1]) : 0;
  std::vector<float> A(M * K), B(N * K), D(M * N, -12345.0f), R(M * N, 0.0f);
  srand(7);
  for (auto& x : A) x = (float)(rand() % 9 - 4);
  for (auto& x : B) x = (float)(rand() % 9 - 4);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += A[m * K + k] * B[n * K + k];
      R[m * N + n] = s;
    }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
  const int smem_bytes = A_BYTES + B_BYTES;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  probe<<<1, 128, smem_bytes>>>(dA, dB, dD, which & 1, (which >> 1) & 1, (which >> 2) & 1, (which >> 3) & 1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("{\"case\": %d, \"error\": \"%s\"}\n", which, cudaGetErrorString(e));
    return 0;
  }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0, first = -1;
  for (int i = 0; i < M * N; ++i)
    if (D[i] != R[i]) {
      if (first < 0) first = i;
      ++bad;
    }
  printf("{\"case\": %d, \"swap_a\": %d, \"swap_b\": %d, \"b_major\": \"%s\", \"b_padded\": %d, \"mismatches\": %d, \"of\": %d", which,
         which & 1, (which >> 1) & 1, (which & 4) ? "K" : "MN", (which >> 3) & 1, bad, M * N);
  if (first >= 0) printf(", \"first\": [%d, %d, %g, %g]", first / N, first % N, D[first], R[first]);
  printf("}\n");
  return 0;
}
