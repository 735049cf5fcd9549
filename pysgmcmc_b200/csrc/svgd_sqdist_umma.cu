// K11 on the tcgen05 tensor cores: squared pairwise distances of the particles through the
// Gram matrix of the CENTRED particles,
//   c_i = x_i - mean(x),   P[i,j] = |c_i|^2 + |c_j|^2 - 2 <c_i, c_j>      (distances do not change)
// (pysgmcmc/samplers/svgd.py:151-152 via tensor_utils.pdist/squareform; oracle/svgd.py).
// Centring keeps |c|^2 of the order of the distances themselves, so the cancellation in the
// formula costs ~1e-6 relative to the particle spread instead of relative to |x|^2.  The dot
// products run as 3xTF32 on the tensor cores (see csrc/svgd_umma.cu for the scheme and
// csrc/umma.cuh for the operand layouts); the FFMA kernel in svgd.cu, which subtracts before
// squaring, stays the implementation for small or unaligned shapes and the test reference.
//
// Tile: 128 particles i x 256 particles j per CTA, contraction over the dimensions in blocks of
// 16, 4-stage ring; both operands are rows of X (K-major as they lie in memory): a quarter-warp
// stores 8 consecutive rows of one 16-byte column = 128 contiguous bytes, conflict-free.
// Only tiles that reach the upper triangle are computed.  A thread of the epilogue owns row i
// and 16 consecutive j: P[i, j] goes out directly for j > i, the mirror image P[j, i] with
// lanes along i (coalesced), the diagonal is written as exact zeros -- P is symmetric bit for
// bit, which K14 relies on when it reads rows of K as columns.
#include "common.cuh"
#include "umma.cuh"

namespace sgmcmc {

constexpr int SQ_BM = 128, SQ_BN = 256, SQ_BK = 16, SQ_STAGES = 4;
constexpr int SQ_PRODUCERS = 256, SQ_THREADS = SQ_PRODUCERS + 32;
constexpr uint32_t SQ_SBO = 128;
constexpr uint32_t SQ_A_LBO = (SQ_BM / 8) * SQ_SBO, SQ_B_LBO = (SQ_BN / 8) * SQ_SBO;      // 2048, 4096
constexpr uint32_t SQ_A_PART = SQ_A_LBO * (SQ_BK / 4), SQ_B_PART = SQ_B_LBO * (SQ_BK / 4);  // 8192, 16384
constexpr uint32_t SQ_STAGE = 2 * SQ_A_PART + 2 * SQ_B_PART;                                // 49152
constexpr uint32_t SQ_SMEM = SQ_STAGES * SQ_STAGE;                                          // 196608
constexpr int SQ_GROUPS = (SQ_BM + SQ_BN) / 8 / (SQ_PRODUCERS / 32);                        // 6 row groups per warp

// column sums over the particles, deterministic: 32 columns x 8 row lanes per CTA
__global__ void __launch_bounds__(256)
svgd_col_mean_kernel(const float* __restrict__ X, float* __restrict__ mean, int n, int D) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int d = blockIdx.x * 32 + tx;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (d < D) {
    int r = ty;
    for (; r + 24 < n; r += 32) {
      s0 += X[(int64_t)r * D + d];
      s1 += X[(int64_t)(r + 8) * D + d];
      s2 += X[(int64_t)(r + 16) * D + d];
      s3 += X[(int64_t)(r + 24) * D + d];
    }
    for (; r < n; r += 8) s0 += X[(int64_t)r * D + d];
  }
  red[ty][tx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (ty == 0 && d < D) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += red[y][tx];
    mean[d] = __fdiv_rn(t, (float)n);
  }
}

// |x_i - mean|^2, one CTA per particle, fixed reduction order
__global__ void __launch_bounds__(256)
svgd_row_norm_kernel(const float* __restrict__ X, const float* __restrict__ mean, float* __restrict__ norms, int D) {
  __shared__ float red[8];
  const float* row = X + (int64_t)blockIdx.x * D;
  float s = 0.f;
  for (int d = threadIdx.x; d < D; d += 256) {
    const float c = __fsub_rn(row[d], mean[d]);
    s = fmaf(c, c, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    norms[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(SQ_THREADS, 1)
svgd_sqdist_umma_kernel(const float* __restrict__ X, const float* __restrict__ mean, const float* __restrict__ norms,
                        float* __restrict__ P, float* __restrict__ partial, int n, int D, int n_slices) {
  const int i0 = blockIdx.y * SQ_BM, j0 = blockIdx.x * SQ_BN;
  if (j0 + SQ_BN - 1 < i0) return;                   // tile entirely below the diagonal (uniform exit)

  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[SQ_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[SQ_STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // blockIdx.z = slice of the contraction (n_slices > 1 when there are too few tiles to fill the SMs): the
  // slice's partial Gram tile goes to `partial`, svgd_sqdist_finalize_kernel adds the slices in a fixed order
  const int nkb_all = (D + SQ_BK - 1) / SQ_BK, kb_per = (nkb_all + n_slices - 1) / n_slices;
  const int kb_begin = (int)blockIdx.z * kb_per;
  const int nkb = min(kb_per, nkb_all - kb_begin);
  const uint32_t smem_base = umma::smem_u32(smem);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < SQ_STAGES; ++s) {
      umma::mbar_init(umma::smem_u32(&full_bar[s]), SQ_PRODUCERS);
      umma::mbar_init(umma::smem_u32(&empty_bar[s]), 1);
    }
    umma::mbar_init(umma::smem_u32(&accum_bar), 1);
    umma::mbar_init_fence();
  }
  if (warp == SQ_PRODUCERS / 32) umma::tmem_alloc<256>(umma::smem_u32(&tmem_slot));
  umma::fence_before_thread_sync();
  __syncthreads();
  umma::fence_after_thread_sync();
  const uint32_t taddr = tmem_slot;

  if (warp < SQ_PRODUCERS / 32) {
    // ------------------------------------------------------------------ producers
    // lane -> (row r of an 8-row group, 16-byte column dq); warp w owns row groups w, w + 8, ...:
    // groups 0..15 are the A rows (particles i0 ..), groups 16..47 the B rows (particles j0 ..)
    const int r = lane & 7, dq = lane >> 3;
    const float* row_ptr[SQ_GROUPS];
    bool row_in[SQ_GROUPS];
    uint32_t s_off[SQ_GROUPS];
#pragma unroll
    for (int e = 0; e < SQ_GROUPS; ++e) {
      const int g = warp + 8 * e;
      const bool is_a = g < SQ_BM / 8;
      const int local = (is_a ? g : g - SQ_BM / 8) * 8 + r;
      const int p = (is_a ? i0 : j0) + local;
      row_in[e] = p < n;
      row_ptr[e] = X + (int64_t)p * D + 4 * dq;
      s_off[e] = (is_a ? 0u : 2 * SQ_A_PART) + (uint32_t)dq * (is_a ? SQ_A_LBO : SQ_B_LBO) + (uint32_t)local * 16;
    }
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    struct Regs {
      float4 v[SQ_GROUPS];
      float4 m;
    };
    auto load = [&](Regs& q, int kb_local) {
      const int kb = kb_begin + kb_local;
      const int d = kb * SQ_BK + 4 * dq;
      const bool din = d < D;
      q.m = din ? __ldg(reinterpret_cast<const float4*>(mean + d)) : zero4;
#pragma unroll
      for (int e = 0; e < SQ_GROUPS; ++e)
        q.v[e] = (din && row_in[e]) ? __ldg(reinterpret_cast<const float4*>(row_ptr[e] + kb * SQ_BK)) : q.m;
      // rows beyond n and columns beyond D load the mean itself: they centre to exact zeros
    };
    auto produce = [&](const Regs& q, int kb) {
      const int s = kb % SQ_STAGES;
      umma::mbar_wait(umma::smem_u32(&empty_bar[s]), (((uint32_t)(kb / SQ_STAGES)) & 1u) ^ 1u);
      uint8_t* stage = smem + (uint32_t)s * SQ_STAGE;
#pragma unroll
      for (int e = 0; e < SQ_GROUPS; ++e) {
        const float4 c = make_float4(__fsub_rn(q.v[e].x, q.m.x), __fsub_rn(q.v[e].y, q.m.y),
                                     __fsub_rn(q.v[e].z, q.m.z), __fsub_rn(q.v[e].w, q.m.w));
        float4 hi, lo;
        umma::split_tf32(c.x, hi.x, lo.x);
        umma::split_tf32(c.y, hi.y, lo.y);
        umma::split_tf32(c.z, hi.z, lo.z);
        umma::split_tf32(c.w, hi.w, lo.w);
        const uint32_t part = (warp + 8 * e) < SQ_BM / 8 ? SQ_A_PART : SQ_B_PART;
        *reinterpret_cast<float4*>(stage + s_off[e]) = hi;
        *reinterpret_cast<float4*>(stage + s_off[e] + part) = lo;
      }
      umma::fence_proxy_async_smem();
      umma::mbar_arrive(umma::smem_u32(&full_bar[s]));
    };

    Regs q0, q1;
    load(q0, 0);
    for (int kb = 0; kb < nkb; kb += 2) {
      if (kb + 1 < nkb) load(q1, kb + 1);
      produce(q0, kb);
      if (kb + 1 < nkb) {
        if (kb + 2 < nkb) load(q0, kb + 2);
        produce(q1, kb + 1);
      }
    }

    // ------------------------------------------------------------------ epilogue
    umma::mbar_wait(umma::smem_u32(&accum_bar), 0);
    umma::fence_after_thread_sync();
    const int qd = warp & 3, half = warp >> 2;
    const int i = i0 + 32 * qd + lane;
    const bool i_in = i < n;
    const float ni = i_in ? norms[i] : 0.0f;
    const uint32_t trow = taddr + ((uint32_t)(32 * qd) << 16);
#pragma unroll 1
    for (int c = 0; c < SQ_BN / 2 / 16; ++c) {
      const int col = (SQ_BN / 2) * half + 16 * c;
      float acc[16];
      umma::tmem_ld16(trow + (uint32_t)col, acc);
      const int jb = j0 + col;
      if (jb + 15 < i0 + 32 * qd || jb >= n) continue;      // warp-uniform: chunk entirely below the diagonal / out of range
      if (n_slices > 1) {
        const int64_t tile = (int64_t)blockIdx.z * gridDim.x * gridDim.y + (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
        float4* dst = reinterpret_cast<float4*>(partial + (tile * SQ_BM + 32 * qd + lane) * SQ_BN + col);
#pragma unroll
        for (int v = 0; v < 4; ++v) dst[v] = make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
        continue;
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int j = jb + e;
        const float nj = j < n ? __ldg(norms + j) : 0.0f;
        // max(.., 0): the rounding of the three terms can leave a tiny negative value for near-coincident particles
        const float p = fmaxf(__fadd_rn(__fadd_rn(ni, nj), __fmul_rn(-2.0f, acc[e])), 0.0f);
        if (i_in && j < n) {
          if (j > i) {
            P[(int64_t)i * n + j] = p;
            P[(int64_t)j * n + i] = p;
          } else if (j == i) {
            P[(int64_t)i * n + i] = 0.0f;
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma::instr_desc_tf32(SQ_BM, SQ_BN, 0, 0);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % SQ_STAGES;
      umma::mbar_wait(umma::smem_u32(&full_bar[s]), (uint32_t)(kb / SQ_STAGES) & 1u);
      umma::fence_after_thread_sync();
      if (lane == 0) {
        const uint32_t stage = smem_base + (uint32_t)s * SQ_STAGE;
#pragma unroll
        for (int ks = 0; ks < SQ_BK / 8; ++ks) {
          const uint32_t a_hi = stage + (uint32_t)ks * 2 * SQ_A_LBO, a_lo = a_hi + SQ_A_PART;
          const uint32_t b_hi = stage + 2 * SQ_A_PART + (uint32_t)ks * 2 * SQ_B_LBO, b_lo = b_hi + SQ_B_PART;
          const uint64_t da_hi = umma::smem_desc(a_hi, SQ_A_LBO, SQ_SBO), da_lo = umma::smem_desc(a_lo, SQ_A_LBO, SQ_SBO);
          const uint64_t db_hi = umma::smem_desc(b_hi, SQ_B_LBO, SQ_SBO), db_lo = umma::smem_desc(b_lo, SQ_B_LBO, SQ_SBO);
          umma::mma_tf32(taddr, da_lo, db_hi, idesc, (kb | ks) != 0);
          umma::mma_tf32(taddr, da_hi, db_lo, idesc, 1);
          umma::mma_tf32(taddr, da_hi, db_hi, idesc, 1);
        }
        umma::commit(umma::smem_u32(&empty_bar[s]));
        if (kb == nkb - 1) umma::commit(umma::smem_u32(&accum_bar));
      }
      __syncwarp();
    }
  }

  umma::fence_before_thread_sync();
  __syncthreads();
  if (warp == SQ_PRODUCERS / 32) {
    umma::fence_after_thread_sync();
    umma::tmem_dealloc<256>(taddr);
  }
}

// Sum of the slices' partial Gram tiles (fixed order) -> distances, same write pattern as the fused epilogue.
// grid (tiles_x, padded rows), 256 threads = the 256 columns of a tile.
__global__ void __launch_bounds__(SQ_BN)
svgd_sqdist_finalize_kernel(const float* __restrict__ partial, const float* __restrict__ norms, float* __restrict__ P,
                            int n, int n_slices, int tiles_y) {
  const int i = blockIdx.y, ty = i / SQ_BM, r = i % SQ_BM;
  const int j0 = blockIdx.x * SQ_BN, j = j0 + threadIdx.x;
  if (j0 + SQ_BN - 1 < ty * SQ_BM || i >= n || j >= n || j < i) return;
  if (j == i) {
    P[(int64_t)i * n + i] = 0.0f;
    return;
  }
  const int64_t tiles = (int64_t)gridDim.x * tiles_y, tile = (int64_t)ty * gridDim.x + blockIdx.x;
  float dot = 0.0f;
  for (int s = 0; s < n_slices; ++s) dot = __fadd_rn(dot, partial[((s * tiles + tile) * SQ_BM + r) * SQ_BN + threadIdx.x]);
  const float p = fmaxf(__fadd_rn(__fadd_rn(norms[i], norms[j]), __fmul_rn(-2.0f, dot)), 0.0f);
  P[(int64_t)i * n + j] = p;
  P[(int64_t)j * n + i] = p;
}

static int sqdist_slices(int n, int D, int n_sm) {
  const int tx = (n + SQ_BN - 1) / SQ_BN, ty = (n + SQ_BM - 1) / SQ_BM;
  int upper = 0;
  for (int y = 0; y < ty; ++y)
    for (int x = 0; x < tx; ++x) upper += (x * SQ_BN + SQ_BN - 1 >= y * SQ_BM);
  const int nkb = (D + SQ_BK - 1) / SQ_BK;
  int s = n_sm / (upper > 0 ? upper : 1);
  s = s < 1 ? 1 : (s > 8 ? 8 : s);
  if (s > nkb / 8) s = nkb / 8 > 0 ? nkb / 8 : 1;          // at least 8 blocks of 16 dimensions per slice
  const int per = (nkb + s - 1) / s;
  return (nkb + per - 1) / per;                             // no empty slice
}

// floats of scratch after the 4096-byte select state: mean[D] + norms[n] (+ padding) + the slices' partial tiles
int64_t svgd_sqdist_work_floats(int n, int D, int n_slices) {
  const int64_t head = ((int64_t)D + n + 3) / 4 * 4;
  const int64_t tiles = (int64_t)((n + SQ_BN - 1) / SQ_BN) * ((n + SQ_BM - 1) / SQ_BM);
  return head + (n_slices > 1 ? (int64_t)n_slices * tiles * SQ_BM * SQ_BN : 0);
}

int svgd_sqdist_best_slices(int n, int D) {
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
  }
  return sqdist_slices(n, D, n_sm);
}

// Requirements (checked by the caller): D % 4 == 0, X 16-byte aligned; work = float[work_floats], 16-byte
// aligned; the contraction is sliced only as far as `work_floats` allows.
int launch_svgd_sqdist_umma(const float* X, float* P, float* work, int64_t work_floats, int n, int D,
                            cudaStream_t stream) {
  {   // per launch, not cached: the attribute belongs to the current device
    const cudaError_t e = cudaFuncSetAttribute(svgd_sqdist_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)SQ_SMEM);
    if (e != cudaSuccess) return set_error(SGMCMC_E_CUDA, "svgd_sqdist_umma_kernel: %s", cudaGetErrorString(e));
  }
  int n_slices = svgd_sqdist_best_slices(n, D);
  while (n_slices > 1 && svgd_sqdist_work_floats(n, D, n_slices) > work_floats) --n_slices;
  if (n_slices > 1) {   // re-balance so that no slice is empty
    const int nkb = (D + SQ_BK - 1) / SQ_BK, per = (nkb + n_slices - 1) / n_slices;
    n_slices = (nkb + per - 1) / per;
  }
  SG_REQUIRE(svgd_sqdist_work_floats(n, D, 1) <= work_floats, SGMCMC_E_INVALID,
             "svgd: scratch too small (%lld floats after the select state, need %lld)", (long long)work_floats,
             (long long)svgd_sqdist_work_floats(n, D, 1));
  float* mean = work;
  float* norms = work + D;
  float* partial = work + ((int64_t)D + n + 3) / 4 * 4;
  svgd_col_mean_kernel<<<(D + 31) / 32, 256, 0, stream>>>(X, mean, n, D);
  if (int rc = check_launch("svgd_col_mean_kernel")) return rc;
  svgd_row_norm_kernel<<<n, 256, 0, stream>>>(X, mean, norms, D);
  if (int rc = check_launch("svgd_row_norm_kernel")) return rc;
  const dim3 grid((unsigned)((n + SQ_BN - 1) / SQ_BN), (unsigned)((n + SQ_BM - 1) / SQ_BM), (unsigned)n_slices);
  svgd_sqdist_umma_kernel<<<grid, SQ_THREADS, SQ_SMEM, stream>>>(X, mean, norms, P, partial, n, D, n_slices);
  if (int rc = check_launch("svgd_sqdist_umma_kernel")) return rc;
  if (n_slices > 1) {
    const dim3 fgrid(grid.x, grid.y * SQ_BM);
    svgd_sqdist_finalize_kernel<<<fgrid, SQ_BN, 0, stream>>>(partial, norms, P, n, n_slices, (int)grid.y);
    return check_launch("svgd_sqdist_finalize_kernel");
  }
  return SGMCMC_OK;
}

}  // namespace sgmcmc
