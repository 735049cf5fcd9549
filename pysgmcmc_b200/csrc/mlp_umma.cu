// Wide layers of K4 for any fully connected architecture (csrc/mlp.cu) on the 5th-generation
// tensor cores (tcgen05 + TMEM): the two GEMMs of a hidden layer whose weight matrix is a genuine
// dense contraction -- BASELINE.json configs[4]'s 1000-512-512 network --
//   forward        H_l^T [n_out x batch]   = W_l^T [n_out x n_in]  H_{l-1}^T [n_in x batch]  (+ b_l, tanh)
//   backward data  dZ_{l-1}^T [n_in x batch] = W_l [n_in x n_out]  dZ_l^T [n_out x batch]    (* (1 - H_{l-1}^2))
// (pysgmcmc/models/bayesian_neural_network.py:28-69 forward, tf.gradients backward).  The chain's
// PRIVATE weight matrix is the M operand (tiles of 128 units), the minibatch the N operand (padded to
// 32 rows), the contraction streams over the other layer width in blocks of 16.  Every weight is read
// from HBM exactly once per pass and used for `batch` FMAs: 10 flop per weight byte, which the FP32
// pipe cannot sustain at HBM speed (the FFMA kernels of mlp.cu run at 0.14 of the copy peak) and the
// tensor pipe can.
//
// fp32 accuracy on the TF32 datapath as in csrc/svgd_umma.cu (3xTF32): operands are split into
// hi = tf32_rn(x), lo = x - hi, each product issued as lo*hi + hi*lo + hi*hi into the fp32 TMEM
// accumulator.  Structure per CTA (one chain x one 128-unit tile), 288 threads:
//   warps 0-7  producers: coalesced global loads of the weight block (3 register buffers ahead),
//              split, stores into a 4-stage shared-memory ring in the no-swizzle K-major canonical
//              layout of the MMA descriptors (forward: W is contiguous along the NON-contracted index,
//              so the producers transpose while storing, conflict-free thanks to 144-byte row groups;
//              backward: 128-bit stores, the core-matrix columns 2336 B apart for the same reason);
//              the activation operand arrives ALREADY split and in canonical order -- the epilogue
//              that produced it wrote it that way -- so its staging is a plain 4 KB copy per block;
//   warp 8     one lane issues 2 k-steps x 3 products of tcgen05.mma (M = 128, N = 32, K = 8),
//              tcgen05.commit releases the stage / publishes the accumulator through mbarriers;
//   epilogue   (warps 0-7, a thread owns one unit = one TMEM lane): bias + tanh (forward) or
//              * (1 - H^2) (backward), then the result leaves twice: plain [unit][BT] for the FFMA
//              kernels (head, weight gradient) and split + canonical for the next tensor-core GEMM.
// Accumulation: tcgen05.mma adds into its TMEM accumulator with truncation, a bias of up to one ulp of
// the running sum per instruction that compounds over the 3 x K / 8 instructions of a long contraction
// (K = 1000: measured 1e-5 relative on the cost with a single accumulator).  So a tile keeps EIGHT
// accumulators of 32 columns: the k-blocks go round-robin into seven of them, which only ever take the
// exact hi*hi products, and the two small cross terms of every block go into the eighth, whose sum --
// and therefore whose ulp -- is 2^-11 of the others'.  The epilogue adds the eight in fp32, round to nearest.
// Shared memory 91 KB per CTA -> 2 CTAs per SM, TMEM 256 columns each (all 512 of the SM).
#include "bnn_common.cuh"
#include "umma.cuh"
#include "mlp_umma.cuh"

namespace sgmcmc {

constexpr int MU_BM = 128, MU_BK = 16;
constexpr int MU_PRODUCERS = 256, MU_THREADS = MU_PRODUCERS + 32;
constexpr uint32_t MU_A_SBO = 144;                      // 8 rows x 16 B, padded (transposing stores hit 32 banks)
constexpr uint32_t MU_A_LBO = 16 * MU_A_SBO + 32;       // 128 rows = 16 groups; +32: the 128-bit stores spread too
constexpr uint32_t MU_A_PART = MU_A_LBO * (MU_BK / 4);  // hi (or lo) of the weight block: 9344 B
constexpr uint32_t MU_B_SBO = 128, MU_B_LBO = 16 * MU_CN, MU_B_PART = MU_B_LBO * (MU_BK / 4);   // 2048 B
constexpr uint32_t MU_STAGE = 2 * MU_A_PART + 2 * MU_B_PART;                                     // 22784 B
constexpr int MU_STAGES = 4;
constexpr uint32_t MU_SMEM = MU_STAGES * MU_STAGE;                                               // 91136 B
constexpr int MU_PF = 3;
constexpr int MU_NACC = 7;                               // hi*hi accumulators (+ 1 for the cross terms): 8 x 32 TMEM columns

struct MuRegs {
  float4 a[2], b;
};

__device__ __forceinline__ float4 ld4_al(const float* p, bool a16) {
  if (a16) return __ldg(reinterpret_cast<const float4*>(p));
  const float2 u = __ldg(reinterpret_cast<const float2*>(p)), v = __ldg(reinterpret_cast<const float2*>(p) + 1);
  return make_float4(u.x, u.y, v.x, v.y);
}

template <bool FWD>
__global__ void __launch_bounds__(MU_THREADS, 2) mlp_gemm_umma_kernel(MuArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[MU_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MU_STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float red[MU_PRODUCERS / 32];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * MU_BM;
  const int64_t chain = blockIdx.y;
  const int M = a.M, K = a.K;
  const int nkb = (K + MU_BK - 1) / MU_BK;
  const uint32_t smem_base = umma::smem_u32(smem);
  const float* __restrict__ W = a.theta + (chain / a.theta_div) * a.D + a.oW;
  float* __restrict__ ws = a.ws + chain * a.ws_floats;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < MU_STAGES; ++s) {
      umma::mbar_init(umma::smem_u32(&full_bar[s]), MU_PRODUCERS);
      umma::mbar_init(umma::smem_u32(&empty_bar[s]), 1);
    }
    umma::mbar_init(umma::smem_u32(&accum_bar), 1);
    umma::mbar_init_fence();
  }
  if (warp == MU_PRODUCERS / 32) umma::tmem_alloc<32 * (MU_NACC + 1)>(umma::smem_u32(&tmem_slot));
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t taddr = tmem_slot;

  if (warp < MU_PRODUCERS / 32) {
    // ------------------------------------------------------------------ producers
    const bool a16 = aligned_to_dev(W, 16);
    // weight block of 16 k x 128 units: two float4 per thread
    int64_t a_goff[2];       // offset of the float4 inside the block (relative to k0 * stride)
    uint32_t a_off[2];
    bool a_in[2];
    int a_k[2];              // FWD: row k of the float4 inside the block (the K edge is per row)
    if (FWD) {
      // W[k][m], m contiguous: lane -> (k within a quad kr, quad of units dql); a warp-load reads 4 rows x 128
      // contiguous bytes, the matching warp-store writes 4 k x 32 units as 32-bit words into 32 distinct banks
      const int kr = lane & 3, dql = lane >> 2;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int tau = 2 * warp + e, kq = tau & 3, dq = 8 * (tau >> 2) + dql;
        a_k[e] = 4 * kq + kr;
        a_in[e] = m0 + 4 * dq < M;
        a_goff[e] = (int64_t)a_k[e] * a.ldw + m0 + 4 * dq;
        a_off[e] = (uint32_t)(dq >> 1) * MU_A_SBO + (uint32_t)(4 * (dq & 1)) * 16 + (uint32_t)kq * MU_A_LBO +
                   (uint32_t)kr * 4;
      }
    } else {
      // W[m][k], k contiguous: a thread moves 4 consecutive k of one unit with 128-bit accesses
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int idx = tid + MU_PRODUCERS * e, m = idx >> 2, kq = idx & 3;
        a_k[e] = 4 * kq;
        a_in[e] = m0 + m < M;
        a_goff[e] = (int64_t)(m0 + m) * a.ldw + 4 * kq;
        a_off[e] = (uint32_t)(m >> 3) * MU_A_SBO + (uint32_t)(m & 7) * 16 + (uint32_t)kq * MU_A_LBO;
      }
    }
    // activation operand: 2 planes x 2 KB per block, already split and in canonical order
    const float* __restrict__ Bsrc = ws + a.oBc + (int64_t)(tid >> 7) * a.b_plane + (tid & 127) * 4;
    const uint32_t b_off = 2 * MU_A_PART + (uint32_t)(tid >> 7) * MU_B_PART + (uint32_t)(tid & 127) * 16;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float sq = 0.0f;

    auto load = [&](MuRegs& r, int kb) {
      const int k0 = kb * MU_BK;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool in = a_in[e] && (k0 + a_k[e] < K);
        r.a[e] = in ? ld4_al(W + (FWD ? (int64_t)k0 * a.ldw : (int64_t)k0) + a_goff[e], a16) : zero4;
      }
      r.b = *reinterpret_cast<const float4*>(Bsrc + (int64_t)kb * (MU_BK / 4) * MU_CN * 4);
    };
    auto produce = [&](const MuRegs& r, int kb) {
      const int s = kb % MU_STAGES;
      const uint32_t parity = ((uint32_t)(kb / MU_STAGES) & 1u) ^ 1u;
      umma::mbar_wait(umma::smem_u32(&empty_bar[s]), parity);
      uint8_t* stage = smem + (uint32_t)s * MU_STAGE;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float4 hi, lo;
        umma::split_tf32(r.a[e].x, hi.x, lo.x);
        umma::split_tf32(r.a[e].y, hi.y, lo.y);
        umma::split_tf32(r.a[e].z, hi.z, lo.z);
        umma::split_tf32(r.a[e].w, hi.w, lo.w);
        if (FWD) {
          sq = fmaf(r.a[e].x, r.a[e].x, sq); sq = fmaf(r.a[e].y, r.a[e].y, sq);
          sq = fmaf(r.a[e].z, r.a[e].z, sq); sq = fmaf(r.a[e].w, r.a[e].w, sq);
          float* ph = reinterpret_cast<float*>(stage + a_off[e]);
          float* pl = reinterpret_cast<float*>(stage + MU_A_PART + a_off[e]);
          ph[0] = hi.x; ph[4] = hi.y; ph[8] = hi.z; ph[12] = hi.w;      // consecutive units = consecutive rows, 16 B apart
          pl[0] = lo.x; pl[4] = lo.y; pl[8] = lo.z; pl[12] = lo.w;
        } else {
          *reinterpret_cast<float4*>(stage + a_off[e]) = hi;
          *reinterpret_cast<float4*>(stage + MU_A_PART + a_off[e]) = lo;
        }
      }
      *reinterpret_cast<float4*>(stage + b_off) = r.b;
      umma::fence_proxy_async_smem();
      umma::mbar_arrive(umma::smem_u32(&full_bar[s]));
    };

    MuRegs r[MU_PF];
#pragma unroll
    for (int p = 0; p < MU_PF - 1; ++p)
      if (p < nkb) load(r[p], p);
    for (int kb0 = 0; kb0 < nkb; kb0 += MU_PF) {
#pragma unroll
      for (int p = 0; p < MU_PF; ++p) {
        const int kb = kb0 + p;
        if (kb < nkb) {
          if (kb + MU_PF - 1 < nkb) load(r[(p + MU_PF - 1) % MU_PF], kb + MU_PF - 1);
          produce(r[p], kb);
        }
      }
    }

    // ------------------------------------------------------------------ epilogue
    umma::mbar_wait(umma::smem_u32(&accum_bar), 0);
    umma::fence_after_thread_sync();
    const int q = warp & 3, hcol = warp >> 2;
    const int u = m0 + 32 * q + lane;                  // this thread's unit = its TMEM lane
    float v[16];
    {
      const uint32_t tbase = taddr + ((uint32_t)(32 * q) << 16) + (uint32_t)(16 * hcol);
      umma::tmem_ld16(tbase + 32 * MU_NACC, v);          // the cross terms first (smallest)
      const int n_main = nkb < MU_NACC ? nkb : MU_NACC;
      for (int j = 0; j < n_main; ++j) {
        float p[16];
        umma::tmem_ld16(tbase + 32 * (uint32_t)j, p);
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = __fadd_rn(v[c], p[c]);
      }
    }
    int64_t row0;
    int rows;
    {
      row0 = a.starts != nullptr ? (int64_t)a.starts[chain] : (chain % a.theta_div) * (int64_t)a.batch;
      const int64_t left = a.n_rows - row0;
      rows = (int)(left < a.batch ? (left < 0 ? 0 : left) : a.batch);
    }
    const int BT = a.bt;
    if (u < M) {
      float bias = 0.0f;
      if (FWD) {
        bias = __ldg(a.theta + (chain / a.theta_div) * a.D + a.ob + u);
        if (hcol == 0) sq = fmaf(bias, bias, sq);
      }
      const float* __restrict__ Hin = ws + a.oHin + (int64_t)u * BT;     // backward: H of this unit
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const int n = 16 * hcol + c;
        if (FWD) v[c] = n < rows ? fast_tanh(v[c] + bias) : 0.0f;
        else {
          const float h = n < BT ? Hin[n] : 0.0f;
          v[c] = n < rows ? v[c] * fmaf(-h, h, 1.0f) : 0.0f;
        }
      }
      float* __restrict__ out = ws + a.oOut + (int64_t)u * BT;            // plain [unit][BT]
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const int n = 16 * hcol + 4 * c4;
        if (n < BT) *reinterpret_cast<float4*>(out + n) = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = 0.0f;
    }
    if (a.oOutc >= 0 && u < a.Mpad) {                   // split + canonical, for the next tensor-core GEMM
      float* __restrict__ oh = ws + a.oOutc + ((int64_t)(u >> 2) * MU_CN + 16 * hcol) * 4 + (u & 3);
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        float hi, lo;
        umma::split_tf32(v[c], hi, lo);
        oh[4 * c] = hi;
        oh[a.out_plane + 4 * c] = lo;
      }
    }
    if (FWD) {                                          // sum of squares of this tile's weights (+ biases): weight prior
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      if (lane == 0) red[warp] = sq;
      asm volatile("bar.sync 1, %0;" ::"n"(MU_PRODUCERS) : "memory");
      if (tid == 0) {
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < MU_PRODUCERS / 32; ++i) s += red[i];
        ws[a.oSq + blockIdx.x] = s;
      }
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer (warp 8)
    constexpr uint32_t idesc = umma::instr_desc_tf32(MU_BM, MU_CN, 0, 0);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % MU_STAGES;
      umma::mbar_wait(umma::smem_u32(&full_bar[s]), (uint32_t)(kb / MU_STAGES) & 1u);
      umma::fence_after_thread_sync();
      if (lane == 0) {
        const uint32_t stage = smem_base + (uint32_t)s * MU_STAGE;
#pragma unroll
        for (int ks = 0; ks < MU_BK / 8; ++ks) {
          const uint32_t a_hi = stage + (uint32_t)ks * 2 * MU_A_LBO, a_lo = a_hi + MU_A_PART;
          const uint32_t b_hi = stage + 2 * MU_A_PART + (uint32_t)ks * 2 * MU_B_LBO, b_lo = b_hi + MU_B_PART;
          const uint64_t da_hi = umma::smem_desc(a_hi, MU_A_LBO, MU_A_SBO), da_lo = umma::smem_desc(a_lo, MU_A_LBO, MU_A_SBO);
          const uint64_t db_hi = umma::smem_desc(b_hi, MU_B_LBO, MU_B_SBO), db_lo = umma::smem_desc(b_lo, MU_B_LBO, MU_B_SBO);
          const uint32_t t_main = taddr + 32u * (uint32_t)(kb % MU_NACC), t_small = taddr + 32u * MU_NACC;
          umma::mma_tf32(t_small, da_lo, db_hi, idesc, (kb | ks) != 0);
          umma::mma_tf32(t_small, da_hi, db_lo, idesc, 1);
          umma::mma_tf32(t_main, da_hi, db_hi, idesc, (kb >= MU_NACC) || ks != 0);
        }
        umma::commit(umma::smem_u32(&empty_bar[s]));                      // stage free once these MMAs retire
        if (kb == nkb - 1) umma::commit(umma::smem_u32(&accum_bar));      // accumulator complete
      }
      __syncwarp();
    }
  }

  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == MU_PRODUCERS / 32) {
    umma::fence_after_thread_sync();
    umma::tmem_dealloc<32 * (MU_NACC + 1)>(taddr);
  }
}

int launch_mlp_gemm_umma(const MuArgs& a, bool fwd, int64_t n_items, cudaStream_t st) {
  SG_REQUIRE(n_items <= 65535, SGMCMC_E_UNSUPPORTED, "mlp (tensor-core layers): at most 65535 chains per launch");
  const dim3 grid((unsigned)((a.M + MU_BM - 1) / MU_BM), (unsigned)n_items);
  if (fwd) {
    cudaFuncSetAttribute(mlp_gemm_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MU_SMEM);
    mlp_gemm_umma_kernel<true><<<grid, MU_THREADS, MU_SMEM, st>>>(a);
  } else {
    cudaFuncSetAttribute(mlp_gemm_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MU_SMEM);
    mlp_gemm_umma_kernel<false><<<grid, MU_THREADS, MU_SMEM, st>>>(a);
  }
  return check_launch("mlp_gemm_umma_kernel");
}

}  // namespace sgmcmc
