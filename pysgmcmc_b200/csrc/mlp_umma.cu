// Wide layers of K4 for any fully connected architecture (csrc/mlp.cu) on the 5th-generation
// tensor cores (tcgen05 + TMEM): the two GEMMs of a hidden layer whose weight matrix is a genuine
// dense contraction -- BASELINE.json configs[4]'s 1000-512-512 network --
//   forward        H_l^T [n_out x batch]   = W_l^T [n_out x n_in]  H_{l-1}^T [n_in x batch]  (+ b_l, tanh)
//   backward data  dZ_{l-1}^T [n_in x batch] = W_l [n_in x n_out]  dZ_l^T [n_out x batch]    (* (1 - H_{l-1}^2))
// (pysgmcmc/models/bayesian_neural_network.py:28-69 forward, tf.gradients backward).  The chain's
// PRIVATE weight matrix is the M operand (tiles of 128 units), the minibatch the N operand (padded to
// 32 rows), the contraction streams over the other layer width in blocks of 32.  Every weight is read
// from HBM exactly once per pass and used for `batch` FMAs: 10 flop per weight byte, which the FP32
// pipe cannot sustain at HBM speed (the FFMA kernels of mlp.cu run at 0.14 of the copy peak) and the
// tensor pipe can.
//
// fp32 accuracy on the TF32 datapath as in csrc/svgd_umma.cu (3xTF32): operands are split into
// hi = tf32_rn(x), lo = x - hi, each product issued as lo*hi + hi*lo + hi*hi into the fp32 TMEM
// accumulator.  Structure per CTA (one chain x one 128-unit tile), 288 threads:
//   warps 0-7  producers: a block of 32 k x 128 units per step.  Forward (W contiguous along the NON-contracted
//              index): a thread loads one quad of units at four consecutive k -- a 4 x 4 piece in registers --
//              and stores it transposed, four consecutive k of one unit per 128-bit store, straight into
//              the no-swizzle K-major canonical layout of the MMA descriptors (144-byte row groups make the
//              eight lanes of a store phase hit eight different 16-byte bank groups).  Backward (W contiguous
//              along k): 128-bit loads and stores, core-matrix columns 2320 B apart for the same reason.
//              The split is Veltkamp's, packed (FMUL2 / FFMA2 / FADD2 per pair).  The first version moved one
//              float4 per thread and block of 16 k with scalar transposing stores and ran at 0.33 of the
//              DRAM peak, issue bound (~60 instructions per float4; cp.async in place of register
//              prefetch changed nothing -- profiles/r02_ncu_mlp_umma_register_prefetch_summary.txt);
//              the activation operand arrives ALREADY split and in canonical order -- the epilogue
//              that produced it wrote it that way -- so its staging is a plain 8 KB copy per block;
//   warp 8     one lane issues 4 k-steps x 3 products of tcgen05.mma (M = 128, N = 32, K = 8),
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
// Shared memory 89 KB per CTA -> 2 CTAs per SM, TMEM 256 columns each (all 512 of the SM).  What bounds it now
// (profiles/r02_ncu_mlp_umma_summary.txt): DRAM 39-45 % busy, 43 % of the stall samples on the first use of a
// loaded weight block -- one block ahead is all the 96 registers of 2 x 288 threads allow (a third buffer
// spills and halves the rate; see profiles/r02_mlp_wide_tcgen05_bk32.jsonl for what else was tried).
#include "bnn_common.cuh"
#include "umma.cuh"
#include "mlp_umma.cuh"

namespace sgmcmc {

constexpr int MU_BM = 128, MU_BK = 32;
constexpr int MU_PRODUCERS = 256, MU_THREADS = MU_PRODUCERS + 32;
constexpr uint32_t MU_A_SBO = 144;                      // 8 rows x 16 B, padded (see the producers)
constexpr uint32_t MU_A_LBO = 16 * MU_A_SBO + 16;       // 128 rows = 16 groups; +16: the backward stores spread too
constexpr uint32_t MU_A_PART = MU_A_LBO * (MU_BK / 4);  // hi (or lo) of the weight block: 18560 B
constexpr uint32_t MU_B_SBO = 128, MU_B_LBO = 16 * MU_CN, MU_B_PART = MU_B_LBO * (MU_BK / 4);   // 4096 B
constexpr uint32_t MU_STAGE = 2 * MU_A_PART + 2 * MU_B_PART;                                     // 45312 B
constexpr int MU_STAGES = 2;
constexpr uint32_t MU_SMEM = MU_STAGES * MU_STAGE;                                               // 90624 B
constexpr int MU_PF = 2;                                 // register buffers of the producers (one block ahead)
constexpr int MU_NACC = 7;                               // hi*hi accumulators (+ 1 for the cross terms): 8 x 32 TMEM columns

struct MuRegs {
  float4 a[4], b[2];
};

// (16-byte aligned source, or 8 for the odd chains of a network whose parameter count is 2 mod 4)
__device__ __forceinline__ float4 ld4_al(const float* p, bool a16) {
  if (a16) return __ldg(reinterpret_cast<const float4*>(p));
  const float2 u = __ldg(reinterpret_cast<const float2*>(p)), v = __ldg(reinterpret_cast<const float2*>(p) + 1);
  return make_float4(u.x, u.y, v.x, v.y);
}

// Veltkamp's splitting of two words at once: p = 8193 x, hi = p - 8192 x (11 significant bits: a TF32 value),
// lo = x - hi (exact) -- FMUL2 + FFMA2 + FADD2
__device__ __forceinline__ void split2(float x0, float x1, float& h0, float& h1, float& l0, float& l1) {
  asm("{\n\t.reg .b64 x, p, h, l, c1, c2;\n\t"
      "mov.b64 x, {%4, %5};\n\t"
      "mov.b64 c1, {%6, %6};\n\t"
      "mov.b64 c2, {%7, %7};\n\t"
      "mul.rn.f32x2 p, x, c1;\n\t"
      "fma.rn.f32x2 h, x, c2, p;\n\t"
      "sub.rn.f32x2 l, x, h;\n\t"
      "mov.b64 {%0, %1}, h;\n\t"
      "mov.b64 {%2, %3}, l;\n\t}"
      : "=f"(h0), "=f"(h1), "=f"(l0), "=f"(l1)
      : "f"(x0), "f"(x1), "f"(8193.0f), "f"(-8192.0f));
}
__device__ __forceinline__ void split4(const float4& v, float4& hi, float4& lo) {
  split2(v.x, v.y, hi.x, hi.y, lo.x, lo.y);
  split2(v.z, v.w, hi.z, hi.w, lo.z, lo.w);
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
  const int m0 = blockIdx.x * MU_BM;                   // the tiles of a chain are neighbours in the grid: they share
  const int64_t chain = a.chain0 + blockIdx.y;         // the chain's activation operand while it is hot in L2
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
    // weight block of 32 k x 128 units: four float4 per thread
    int64_t a_goff[4];       // offsets of the float4s inside the block
    uint32_t a_off[4];       // FWD: store of unit 4 dq + i;  BWD: store of float4 j
    bool a_in[4];
    int kq;
    if (FWD) {
      // W[k][m], m contiguous: warp = quad of k (rows k0 + 4 kq + j), lane = quad of units dq: a warp-load
      // reads 512 contiguous bytes of one row; the 4 x 4 piece leaves as 4 k of one unit per store
      kq = warp;
      const int dq = lane;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a_in[j] = m0 + 4 * dq < M;
        a_goff[j] = (int64_t)(4 * kq + j) * a.ldw + m0 + 4 * dq;
        a_off[j] = (uint32_t)(dq >> 1) * MU_A_SBO + (uint32_t)(4 * (dq & 1) + j) * 16 + (uint32_t)kq * MU_A_LBO;
      }
    } else {
      // W[m][k], k contiguous: lane = (row within a group of 4, quad of k): a warp-load reads 4 rows x 128 bytes
      kq = lane & 7;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = 16 * warp + 4 * j + (lane >> 3);
        a_in[j] = m0 + m < M;
        a_goff[j] = (int64_t)(m0 + m) * a.ldw + 4 * kq;
        a_off[j] = (uint32_t)(m >> 3) * MU_A_SBO + (uint32_t)(m & 7) * 16 + (uint32_t)kq * MU_A_LBO;
      }
    }
    // activation operand: 2 planes x 4 KB per block, already split and in canonical order
    const float* Bsrc[2];
    uint32_t b_off[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int idx = tid + MU_PRODUCERS * e;
      Bsrc[e] = ws + a.oBc + (int64_t)(idx >> 8) * a.b_plane + (idx & 255) * 4;
      b_off[e] = 2 * MU_A_PART + (uint32_t)(idx >> 8) * MU_B_PART + (uint32_t)(idx & 255) * 16;
    }
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float sq = 0.0f, sq1 = 0.0f;

    auto load = [&](MuRegs& r, int kb) {
      const int k0 = kb * MU_BK;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool in = a_in[j] && (k0 + 4 * kq + (FWD ? j : 0) < K);
        r.a[j] = in ? ld4_al(W + (FWD ? (int64_t)k0 * a.ldw : (int64_t)k0) + a_goff[j], a16) : zero4;
      }
#pragma unroll
      for (int e = 0; e < 2; ++e)
        r.b[e] = *reinterpret_cast<const float4*>(Bsrc[e] + (int64_t)kb * (MU_BK / 4) * MU_CN * 4);
    };
    auto produce = [&](const MuRegs& r, int kb) {
      const int s = kb % MU_STAGES;
      const uint32_t parity = ((uint32_t)(kb / MU_STAGES) & 1u) ^ 1u;
      float4 hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) split4(r.a[j], hi[j], lo[j]);
      if (FWD) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          sq = fmaf(r.a[j].x, r.a[j].x, sq); sq1 = fmaf(r.a[j].y, r.a[j].y, sq1);
          sq = fmaf(r.a[j].z, r.a[j].z, sq); sq1 = fmaf(r.a[j].w, r.a[j].w, sq1);
        }
      }
      umma::mbar_wait(umma::smem_u32(&empty_bar[s]), parity);
      uint8_t* stage = smem + (uint32_t)s * MU_STAGE;
      if (FWD) {                                        // transposed: unit 4 dq + i gets its 4 consecutive k
        *reinterpret_cast<float4*>(stage + a_off[0]) = make_float4(hi[0].x, hi[1].x, hi[2].x, hi[3].x);
        *reinterpret_cast<float4*>(stage + a_off[1]) = make_float4(hi[0].y, hi[1].y, hi[2].y, hi[3].y);
        *reinterpret_cast<float4*>(stage + a_off[2]) = make_float4(hi[0].z, hi[1].z, hi[2].z, hi[3].z);
        *reinterpret_cast<float4*>(stage + a_off[3]) = make_float4(hi[0].w, hi[1].w, hi[2].w, hi[3].w);
        *reinterpret_cast<float4*>(stage + MU_A_PART + a_off[0]) = make_float4(lo[0].x, lo[1].x, lo[2].x, lo[3].x);
        *reinterpret_cast<float4*>(stage + MU_A_PART + a_off[1]) = make_float4(lo[0].y, lo[1].y, lo[2].y, lo[3].y);
        *reinterpret_cast<float4*>(stage + MU_A_PART + a_off[2]) = make_float4(lo[0].z, lo[1].z, lo[2].z, lo[3].z);
        *reinterpret_cast<float4*>(stage + MU_A_PART + a_off[3]) = make_float4(lo[0].w, lo[1].w, lo[2].w, lo[3].w);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          *reinterpret_cast<float4*>(stage + a_off[j]) = hi[j];
          *reinterpret_cast<float4*>(stage + MU_A_PART + a_off[j]) = lo[j];
        }
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) *reinterpret_cast<float4*>(stage + b_off[e]) = r.b[e];
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
    sq += sq1;

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

int launch_mlp_gemm_umma(const MuArgs& a_in, bool fwd, int64_t n_items, cudaStream_t st) {
  MuArgs a = a_in;
  SG_REQUIRE((a.M + MU_BM - 1) / MU_BM <= 65535 * 32, SGMCMC_E_UNSUPPORTED, "mlp (tensor-core layers): layer too wide");
  if (fwd) cudaFuncSetAttribute(mlp_gemm_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MU_SMEM);
  else cudaFuncSetAttribute(mlp_gemm_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MU_SMEM);
  for (int64_t c0 = 0; c0 < n_items; c0 += 65535) {      // grid.y holds the chains: 65535 per launch
    const int64_t nc = n_items - c0 < 65535 ? n_items - c0 : 65535;
    a.chain0 = c0;
    const dim3 grid((unsigned)((a.M + MU_BM - 1) / MU_BM), (unsigned)nc);
    if (fwd) mlp_gemm_umma_kernel<true><<<grid, MU_THREADS, MU_SMEM, st>>>(a);
    else mlp_gemm_umma_kernel<false><<<grid, MU_THREADS, MU_SMEM, st>>>(a);
    if (int rc = check_launch("mlp_gemm_umma_kernel")) return rc;
  }
  return SGMCMC_OK;
}

}  // namespace sgmcmc
