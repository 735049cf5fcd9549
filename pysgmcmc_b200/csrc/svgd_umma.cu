// K14 on the 5th-generation tensor cores (tcgen05 + TMEM): the Stein-direction GEMM
//   [KG | KX] = K[n x n] @ [G | X][n x 2D]           (pysgmcmc/samplers/svgd.py:130-133,162-166)
// with the AdaGrad history and the particle update in the epilogue -- the same contract as
// svgd_update_kernel in svgd.cu (the FFMA implementation, kept for small or unaligned shapes
// and as the second implementation the tests compare against).
//
// fp32 accuracy on a TF32 datapath: every operand is split on the CUDA cores into
// hi = tf32(x) (round to nearest) and lo = x - hi, and each product is issued three times,
// lo*hi + hi*lo + hi*hi, accumulating in the fp32 TMEM accumulator ("3xTF32", error ~2^-21 per
// product instead of TF32's 2^-11).  That split is why there is no TMA here: the operands
// have to pass through registers anyway, so the producer warps load fp32 tiles with coalesced
// 128-bit loads, split them and write hi and lo straight into the no-swizzle canonical layouts
// the MMA descriptors describe (csrc/umma.cuh; conventions pinned by tools/micro/umma_probe.cu).
//
// CTA = MT x 128 particles (rows i) x 128 dimensions (d); UMMA tile M = 128, N = 256: columns
// 0..127 of an accumulator are KG, 128..255 are KX, so K is read once for both products.  MT = 2
// row tiles per CTA share one B tile and fill all 512 TMEM columns: twice the tensor work for
// 1.33x the operand staging, which is what bounds the kernel (151-165 against 109-125 TFLOP/s).
// Contraction over the particles j in blocks of 16, a 4-stage (MT = 2: 3-stage) shared-memory ring:
//   warps 0-7  producers: global -> registers (one block ahead) -> split -> shared; then epilogue
//   warp  8    one lane issues 6 tcgen05.mma per block (2 k-steps x 3 products), tcgen05.commit
//              releases the stage / publishes the accumulator through mbarriers
// Shared memory per stage: A hi+lo 2 x 9 KB per row tile, B hi+lo 2 x 18 KB, both K-major with the
// 8-row groups padded to 144 B so that the scalar transposing stores of a warp -- 4 k x 8 column
// quads per instruction -- hit 32 distinct banks (A is read as rows of the symmetric K, so its global
// loads are as coalesced as B's: the first version gathered one row of K per lane and spent a third
// of the LSU data-pipe wavefronts on that).  TMEM: 256 columns per row tile.
#include "common.cuh"
#include "umma.cuh"

namespace sgmcmc {

constexpr int UM_BM = 128, UM_BN = 128, UM_BK = 16;
constexpr int UM_PRODUCERS = 256, UM_THREADS = UM_PRODUCERS + 32;
constexpr uint32_t UM_A_SBO = 144, UM_A_LBO = 16 * UM_A_SBO;          // 128 rows = 16 groups (padded like B)
constexpr uint32_t UM_B_SBO = 144, UM_B_LBO = 32 * UM_B_SBO;          // 256 rows = 32 groups
constexpr uint32_t UM_A_PART = UM_A_LBO * (UM_BK / 4);                // hi (or lo) of A: 9216
constexpr uint32_t UM_B_PART = UM_B_LBO * (UM_BK / 4);                // hi (or lo) of B: 18432
constexpr uint32_t UM_X_ROWS_OFF = 16 * UM_B_SBO;                     // B rows 128..255 (the X half)
// MT = row tiles (of 128 particles) per CTA sharing one B tile: MT = 2 fills all 512 TMEM columns and does
// twice the tensor work for 1.33x the staging work (the kernel is bound by the staging)
template <int MT> struct UmCfg {
  static constexpr int STAGES = MT == 1 ? 4 : 3;
  static constexpr uint32_t STAGE = 2 * MT * UM_A_PART + 2 * UM_B_PART;   // 55296 / 73728
  static constexpr uint32_t SMEM = STAGES * STAGE;                         // 221184 / 221184
  static constexpr int TMEM_COLS = 256 * MT;
};

template <int MT>
struct UmRegs {
  float4 a[MT][2], g0, x0, g1, x1;
};

template <int PF, int MT>     // PF: register buffers of the producers (global loads run PF - 1 blocks ahead)
__global__ void __launch_bounds__(UM_THREADS, 1)
svgd_update_umma_kernel(const float* __restrict__ K, const float* __restrict__ X, const float* __restrict__ G,
                        const float* __restrict__ ksum, const float* __restrict__ bw, float* __restrict__ hist,
                        float* __restrict__ Xout, int n, int D, float eps, float alpha, float one_minus_alpha,
                        float fudge) {
  constexpr int UM_STAGES = UmCfg<MT>::STAGES;
  constexpr uint32_t UM_STAGE = UmCfg<MT>::STAGE;
  constexpr uint32_t UM_A_ALL = MT * UM_A_PART;      // all hi (or all lo) A tiles of a stage
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[UM_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[UM_STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int d0 = blockIdx.x * UM_BN, i0 = blockIdx.y * (UM_BM * MT);
  const int nkb = (n + UM_BK - 1) / UM_BK;
  const uint32_t smem_base = umma::smem_u32(smem);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < UM_STAGES; ++s) {
      umma::mbar_init(umma::smem_u32(&full_bar[s]), UM_PRODUCERS);
      umma::mbar_init(umma::smem_u32(&empty_bar[s]), 1);
    }
    umma::mbar_init(umma::smem_u32(&accum_bar), 1);
    umma::mbar_init_fence();
  }
  if (warp == UM_PRODUCERS / 32) umma::tmem_alloc<UmCfg<MT>::TMEM_COLS>(umma::smem_u32(&tmem_slot));
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t taddr = tmem_slot;

  if (warp < UM_PRODUCERS / 32) {
    // ------------------------------------------------------------------ producers
    // Both operands are read along their contiguous index with the contraction index j as the row:
    // B[j, d] = G / X rows, A[i, j] = K[j, i] (K is symmetric bit for bit, see K11).
    // lane -> (k within a quad kr, quad of columns dql); a warp-load reads 4 rows x 128 contiguous bytes,
    // the matching warp-store writes 4 k x 32 columns as 32-bit words into 32 distinct banks.
    const int kr = lane & 3, dql = lane >> 2;
    int bk[2];
    bool b_col_in[2];
    int64_t b_goff[2];
    uint32_t b_off[2];
    bool a_col_in[MT][2];
    int64_t a_goff[MT][2];
    uint32_t a_off[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int tau = 2 * warp + e, kq = tau & 3, dq = 8 * (tau >> 2) + dql;
      bk[e] = 4 * kq + kr;
      const int d = d0 + 4 * dq;
      b_col_in[e] = d < D;
      b_goff[e] = (int64_t)bk[e] * D + d;
      b_off[e] = (uint32_t)(dq >> 1) * UM_B_SBO + (uint32_t)(4 * (dq & 1)) * 16 + (uint32_t)kq * UM_B_LBO +
                 (uint32_t)kr * 4;
      a_off[e] = (uint32_t)(dq >> 1) * UM_A_SBO + (uint32_t)(4 * (dq & 1)) * 16 + (uint32_t)kq * UM_A_LBO +
                 (uint32_t)kr * 4;
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        const int i = i0 + t * UM_BM + 4 * dq;
        a_col_in[t][e] = i < n;
        a_goff[t][e] = (int64_t)bk[e] * n + i;
      }
    }
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    auto load = [&](UmRegs<MT>& r, int kb) {
      const int j0 = kb * UM_BK;
      const bool k_in0 = (j0 + bk[0]) < n, k_in1 = (j0 + bk[1]) < n;
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        r.a[t][0] = (k_in0 && a_col_in[t][0]) ? __ldg(reinterpret_cast<const float4*>(K + (int64_t)j0 * n + a_goff[t][0])) : zero4;
        r.a[t][1] = (k_in1 && a_col_in[t][1]) ? __ldg(reinterpret_cast<const float4*>(K + (int64_t)j0 * n + a_goff[t][1])) : zero4;
      }
      const bool in0 = b_col_in[0] && k_in0, in1 = b_col_in[1] && k_in1;
      const int64_t o0 = (int64_t)j0 * D + b_goff[0], o1 = (int64_t)j0 * D + b_goff[1];
      r.g0 = in0 ? __ldg(reinterpret_cast<const float4*>(G + o0)) : zero4;
      r.x0 = in0 ? __ldg(reinterpret_cast<const float4*>(X + o0)) : zero4;
      r.g1 = in1 ? __ldg(reinterpret_cast<const float4*>(G + o1)) : zero4;
      r.x1 = in1 ? __ldg(reinterpret_cast<const float4*>(X + o1)) : zero4;
    };

    auto split4 = [](const float4& v, float4& hi, float4& lo) {
      umma::split_tf32(v.x, hi.x, lo.x);
      umma::split_tf32(v.y, hi.y, lo.y);
      umma::split_tf32(v.z, hi.z, lo.z);
      umma::split_tf32(v.w, hi.w, lo.w);
    };

    auto store_b = [&](uint8_t* b_hi, uint32_t off, const float4& v, uint32_t lo_off) {
      float4 hi, lo;
      split4(v, hi, lo);
      float* ph = reinterpret_cast<float*>(b_hi + off);
      float* pl = reinterpret_cast<float*>(b_hi + lo_off + off);
      ph[0] = hi.x; ph[4] = hi.y; ph[8] = hi.z; ph[12] = hi.w;      // consecutive d = consecutive rows, 16 B apart
      pl[0] = lo.x; pl[4] = lo.y; pl[8] = lo.z; pl[12] = lo.w;
    };

    auto produce = [&](const UmRegs<MT>& r, int kb) {
      const int s = kb % UM_STAGES;
      const uint32_t parity = ((uint32_t)(kb / UM_STAGES) & 1u) ^ 1u;
      umma::mbar_wait(umma::smem_u32(&empty_bar[s]), parity);
      uint8_t* stage = smem + (uint32_t)s * UM_STAGE;
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        store_b(stage + (uint32_t)t * UM_A_PART, a_off[0], r.a[t][0], UM_A_ALL);
        store_b(stage + (uint32_t)t * UM_A_PART, a_off[1], r.a[t][1], UM_A_ALL);
      }
      uint8_t* b_hi = stage + 2 * UM_A_ALL;
      store_b(b_hi, b_off[0], r.g0, UM_B_PART);
      store_b(b_hi, b_off[0] + UM_X_ROWS_OFF, r.x0, UM_B_PART);
      store_b(b_hi, b_off[1], r.g1, UM_B_PART);
      store_b(b_hi, b_off[1] + UM_X_ROWS_OFF, r.x1, UM_B_PART);
      umma::fence_proxy_async_smem();
      umma::mbar_arrive(umma::smem_u32(&full_bar[s]));
    };

    UmRegs<MT> r[PF];
#pragma unroll
    for (int p = 0; p < PF - 1; ++p)
      if (p < nkb) load(r[p], p);
    for (int kb0 = 0; kb0 < nkb; kb0 += PF) {
#pragma unroll
      for (int p = 0; p < PF; ++p) {
        const int kb = kb0 + p;
        if (kb < nkb) {
          if (kb + PF - 1 < nkb) load(r[(p + PF - 1) % PF], kb + PF - 1);
          produce(r[p], kb);
        }
      }
    }

    // ------------------------------------------------------------------ epilogue
    umma::mbar_wait(umma::smem_u32(&accum_bar), 0);
    umma::fence_after_thread_sync();
    const int q = warp & 3, hcol = warp >> 2;
    const float h2 = bw[2], nf = (float)n;
#pragma unroll 1
    for (int tc = 0; tc < MT * 4; ++tc) {
      const int t = tc >> 2, c = tc & 3;
      const int i = i0 + t * UM_BM + 32 * q + lane;
      const bool row_in = i < n;
      const float ks = row_in ? ksum[i] : 0.0f;
      const uint32_t trow = taddr + ((uint32_t)(32 * q) << 16) + (uint32_t)(t * 2 * UM_BN);
      const int col = 64 * hcol + 16 * c;
      float accg[16], accx[16];
      umma::tmem_ld16(trow + (uint32_t)col, accg);
      umma::tmem_ld16(trow + (uint32_t)(UM_BN + col), accx);
      if (!row_in) continue;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int d = d0 + col + 4 * v;
        if (d >= D) continue;
        const int64_t off = (int64_t)i * D + d;
        const float4 x4 = *reinterpret_cast<const float4*>(X + off);
        const float4 h4 = *reinterpret_cast<const float4*>(hist + off);
        const float xv[4] = {x4.x, x4.y, x4.z, x4.w}, hv[4] = {h4.x, h4.y, h4.z, h4.w};
        float xo[4], ho[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float kgrad = __fdiv_rn(__fadd_rn(-accx[4 * v + e], __fmul_rn(xv[e], ks)), h2);
          const float phi = __fdiv_rn(__fadd_rn(accg[4 * v + e], kgrad), nf);
          ho[e] = __fadd_rn(__fmul_rn(alpha, hv[e]), __fmul_rn(one_minus_alpha, __fmul_rn(phi, phi)));
          const float adj = __fdiv_rn(phi, __fadd_rn(fudge, __fsqrt_rn(ho[e])));
          xo[e] = __fsub_rn(xv[e], __fmul_rn(eps, adj));
        }
        *reinterpret_cast<float4*>(hist + off) = make_float4(ho[0], ho[1], ho[2], ho[3]);
        *reinterpret_cast<float4*>(Xout + off) = make_float4(xo[0], xo[1], xo[2], xo[3]);
      }
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer (warp 8)
    constexpr uint32_t idesc = umma::instr_desc_tf32(UM_BM, 2 * UM_BN, 0, 0);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % UM_STAGES;
      umma::mbar_wait(umma::smem_u32(&full_bar[s]), (uint32_t)(kb / UM_STAGES) & 1u);
      umma::fence_after_thread_sync();
      if (lane == 0) {
        const uint32_t stage = smem_base + (uint32_t)s * UM_STAGE;
#pragma unroll
        for (int ks = 0; ks < UM_BK / 8; ++ks) {
          const uint32_t b_hi = stage + 2 * UM_A_ALL + (uint32_t)ks * 2 * UM_B_LBO, b_lo = b_hi + UM_B_PART;
          const uint64_t db_hi = umma::smem_desc(b_hi, UM_B_LBO, UM_B_SBO), db_lo = umma::smem_desc(b_lo, UM_B_LBO, UM_B_SBO);
#pragma unroll
          for (int t = 0; t < MT; ++t) {
            const uint32_t a_hi = stage + (uint32_t)t * UM_A_PART + (uint32_t)ks * 2 * UM_A_LBO, a_lo = a_hi + UM_A_ALL;
            const uint64_t da_hi = umma::smem_desc(a_hi, UM_A_LBO, UM_A_SBO), da_lo = umma::smem_desc(a_lo, UM_A_LBO, UM_A_SBO);
            const uint32_t td = taddr + (uint32_t)(t * 2 * UM_BN);
            umma::mma_tf32(td, da_lo, db_hi, idesc, (kb | ks) != 0);      // small terms first
            umma::mma_tf32(td, da_hi, db_lo, idesc, 1);
            umma::mma_tf32(td, da_hi, db_hi, idesc, 1);
          }
        }
        umma::commit(umma::smem_u32(&empty_bar[s]));                      // stage free once these MMAs retire
        if (kb == nkb - 1) umma::commit(umma::smem_u32(&accum_bar));      // accumulator complete
      }
      __syncwarp();
    }
  }

  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == UM_PRODUCERS / 32) {
    umma::fence_after_thread_sync();
    umma::tmem_dealloc<UmCfg<MT>::TMEM_COLS>(taddr);
  }
}

// Launch helper used by sgmcmc_svgd_update_f32 (svgd.cu).  Requirements (checked by the caller):
// n % 4 == 0, D % 4 == 0, all pointers 16-byte aligned.
template <int PF, int MT>
static int launch_umma_pf(const float* K, const float* X, const float* G, const float* ksum, const float* bw,
                          float* hist, float* Xout, int n, int D, float eps, float alpha, float one_minus_alpha,
                          float fudge, cudaStream_t stream) {
  {   // per launch, not cached: the attribute belongs to the current device
    const cudaError_t e = cudaFuncSetAttribute(svgd_update_umma_kernel<PF, MT>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UmCfg<MT>::SMEM);
    if (e != cudaSuccess) return set_error(SGMCMC_E_CUDA, "svgd_update_umma_kernel: %s", cudaGetErrorString(e));
  }
  const dim3 grid((unsigned)((D + UM_BN - 1) / UM_BN), (unsigned)((n + UM_BM * MT - 1) / (UM_BM * MT)));
  SG_REQUIRE(grid.y <= 65535, SGMCMC_E_UNSUPPORTED, "svgd: too many particles");
  svgd_update_umma_kernel<PF, MT><<<grid, UM_THREADS, UmCfg<MT>::SMEM, stream>>>(
      K, X, G, ksum, bw, hist, Xout, n, D, eps, alpha, one_minus_alpha, fudge);
  return check_launch("svgd_update_umma_kernel");
}

// prefetch: register buffers of the producers (2..4); 0 = the default
int launch_svgd_update_umma(const float* K, const float* X, const float* G, const float* ksum, const float* bw,
                            float* hist, float* Xout, int n, int D, float eps, float alpha, float one_minus_alpha,
                            float fudge, int prefetch, cudaStream_t stream) {
  // measured (profiles/r01_svgd_k14_prefetch_sweep.jsonl): 2, 3 and 4 buffers run within 2 % of each other -- the
  // kernel is bound by the LSU data pipe (global loads + staging stores), not by load latency -- so 2 is the default
  // variant 3: one 128-row tile per CTA, 3 buffers; 4: one tile, 2 buffers (the round-1 first version);
  // default: two row tiles per CTA (all 512 TMEM columns) whenever there are more than 128 particles
  switch (prefetch) {
    case 3: return launch_umma_pf<3, 1>(K, X, G, ksum, bw, hist, Xout, n, D, eps, alpha, one_minus_alpha, fudge, stream);
    case 4: return launch_umma_pf<2, 1>(K, X, G, ksum, bw, hist, Xout, n, D, eps, alpha, one_minus_alpha, fudge, stream);
    default: {
      // one CTA per SM (shared memory), so time ~ waves x work per CTA; a two-tile CTA costs ~1.45-1.55x a
      // one-tile CTA (measured: profiles/r01_svgd_k14_prefetch_sweep.jsonl, 1024 / 4096 / 8192 particles)
      static int n_sm = 0;
      if (n_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
      }
      const long long cols = (D + UM_BN - 1) / UM_BN;
      const long long c1 = cols * ((n + UM_BM - 1) / UM_BM), c2 = cols * ((n + 2 * UM_BM - 1) / (2 * UM_BM));
      const double cost1 = (double)((c1 + n_sm - 1) / n_sm), cost2 = 1.55 * (double)((c2 + n_sm - 1) / n_sm);
      if (n > UM_BM && cost2 < cost1)
        return launch_umma_pf<2, 2>(K, X, G, ksum, bw, hist, Xout, n, D, eps, alpha, one_minus_alpha, fudge, stream);
      return launch_umma_pf<2, 1>(K, X, G, ksum, bw, hist, Xout, n, D, eps, alpha, one_minus_alpha, fudge, stream);
    }
  }
}

}  // namespace sgmcmc
