// K8: chain-moment and lagged-variogram reductions for the convergence diagnostics
// (Gelman-Rubin R-hat and variogram ESS; formulas in
// pysgmcmc/diagnostics/sampler_diagnostics.py:76-82,153-161, which the reference
// delegates to pymc3; restated in oracle/diagnostics.py).
//
// trace is [n_draws, C, D] fp32 (d fastest).  A CTA owns a tile of 32 dimensions x
// CHAINS_PER_BLOCK chains: threadIdx.x walks the dimensions (coalesced 128 B rows),
// threadIdx.y strides over the tile's chains; per-(chain, dim) sums run in fp64 on data
// shifted by the first draw, the tile is reduced over chains in shared memory and one fp64
// atomicAdd per (CTA, dim) lands in the output.  The outputs are per-dimension SUMS over the
// local chains -- exactly the quantities one NCCL all-reduce combines across GPUs (K9).
// HBM bound: one pass over the trace for the moments, one pass per lag for the variogram.
#include "common.cuh"

namespace sgmcmc {

constexpr int MT_DX = 32;                // dimensions per CTA
constexpr int MT_DY = 8;                 // chain lanes per CTA
constexpr int CHAINS_PER_BLOCK = 256;

__global__ void __launch_bounds__(MT_DX * MT_DY)
chain_moments_kernel(const float* __restrict__ trace, double* __restrict__ sums, int64_t n_draws,
                     int64_t n_chains, int64_t n_dims) {
  __shared__ double red[3][MT_DY][MT_DX];
  const int64_t d = (int64_t)blockIdx.x * MT_DX + threadIdx.x;
  const int64_t c0 = (int64_t)blockIdx.y * CHAINS_PER_BLOCK;
  const int64_t c1 = min(c0 + CHAINS_PER_BLOCK, n_chains);
  double s_mean = 0.0, s_mean2 = 0.0, s_var = 0.0;
  if (d < n_dims) {
    const int64_t stride = n_chains * n_dims;
    for (int64_t j = c0 + threadIdx.y; j < c1; j += MT_DY) {
      const float* p = trace + j * n_dims + d;
      const double x0 = (double)p[0];
      // four independent accumulator pairs keep four strided loads in flight per thread
      double a1[4] = {0.0, 0.0, 0.0, 0.0}, a2[4] = {0.0, 0.0, 0.0, 0.0};
      int64_t i = 1;
      for (; i + 3 < n_draws; i += 4) {
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = __ldcs(p + (i + q) * stride);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const double x = (double)v[q] - x0;
          a1[q] += x;
          a2[q] += x * x;
        }
      }
      for (; i < n_draws; ++i) {
        const double x = (double)p[i * stride] - x0;
        a1[0] += x;
        a2[0] += x * x;
      }
      const double s1 = (a1[0] + a1[1]) + (a1[2] + a1[3]);
      const double s2 = (a2[0] + a2[1]) + (a2[2] + a2[3]);
      const double n = (double)n_draws;
      const double mean = x0 + s1 / n;
      const double var = n_draws > 1 ? (s2 - s1 * s1 / n) / (n - 1.0) : 0.0;
      s_mean += mean;
      s_mean2 += mean * mean;
      s_var += var;
    }
  }
  red[0][threadIdx.y][threadIdx.x] = s_mean;
  red[1][threadIdx.y][threadIdx.x] = s_mean2;
  red[2][threadIdx.y][threadIdx.x] = s_var;
  __syncthreads();
  if (threadIdx.y < 3 && d < n_dims) {
    double t = 0.0;
#pragma unroll
    for (int y = 0; y < MT_DY; ++y) t += red[threadIdx.y][y][threadIdx.x];
    atomicAdd(sums + (int64_t)threadIdx.y * n_dims + d, t);
  }
}

__global__ void __launch_bounds__(MT_DX * MT_DY)
variogram_kernel(const float* __restrict__ trace, double* __restrict__ out, int64_t n_draws,
                 int64_t n_chains, int64_t n_dims, int64_t lag0) {
  __shared__ double red[MT_DY][MT_DX];
  const int64_t d = (int64_t)blockIdx.x * MT_DX + threadIdx.x;
  const int64_t c0 = (int64_t)blockIdx.y * CHAINS_PER_BLOCK;
  const int64_t c1 = min(c0 + CHAINS_PER_BLOCK, n_chains);
  const int64_t t = lag0 + blockIdx.z;
  double acc = 0.0;
  if (d < n_dims && t < n_draws) {
    const int64_t stride = n_chains * n_dims;
    for (int64_t j = c0 + threadIdx.y; j < c1; j += MT_DY) {
      const float* p = trace + j * n_dims + d;
      for (int64_t i = t; i < n_draws; ++i) {
        const double diff = (double)p[i * stride] - (double)p[(i - t) * stride];
        acc += diff * diff;
      }
    }
  }
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && d < n_dims) {
    double s = 0.0;
#pragma unroll
    for (int y = 0; y < MT_DY; ++y) s += red[y][threadIdx.x];
    atomicAdd(out + (int64_t)blockIdx.z * n_dims + d, s);
  }
}


// ---- vectorised paths ----------------------------------------------------------------------
// The scalar kernels above touch 128 B per (chain, draw) and warp, and consecutive draws of a
// chain are n_chains * D * 4 bytes apart: every request opens its own DRAM page (measured
// 1.3-1.5 TB/s).  Here a thread owns 4 (moments) or 2 (variogram) consecutive dimensions, a CTA
// reads 4 KB / 1 KB contiguous per (chain, draw) row, and every thread keeps 8-16 independent
// 128- / 64-bit loads in flight.
constexpr int MV_THREADS = 256;
constexpr int MV_CHAINS = 32;            // chains per CTA
constexpr int MV_UNROLL = 8;

__global__ void __launch_bounds__(MV_THREADS, 2)
chain_moments_vec_kernel(const float4* __restrict__ trace4, double* __restrict__ sums, int64_t n_draws,
                         int64_t n_chains, int64_t d4) {
  const int64_t q = (int64_t)blockIdx.x * MV_THREADS + threadIdx.x;
  if (q >= d4) return;
  const int64_t c0 = (int64_t)blockIdx.y * MV_CHAINS;
  const int64_t c1 = min(c0 + MV_CHAINS, n_chains);
  const int64_t stride = n_chains * d4;
  const double n = (double)n_draws;
  double s_mean[4] = {0.0, 0.0, 0.0, 0.0}, s_mean2[4] = {0.0, 0.0, 0.0, 0.0}, s_var[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t j = c0; j < c1; ++j) {
    const float4* p = trace4 + j * d4 + q;
    const float4 f0 = __ldcs(p);
    const double x0[4] = {(double)f0.x, (double)f0.y, (double)f0.z, (double)f0.w};
    double a1[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
    double a2[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
    for (int64_t i = 1; i < n_draws; i += MV_UNROLL) {
      float4 v[MV_UNROLL];
#pragma unroll
      for (int u = 0; u < MV_UNROLL; ++u)      // past the end: the first draw again, which adds exactly 0
        v[u] = i + u < n_draws ? __ldcs(p + (i + u) * stride) : f0;
#pragma unroll
      for (int u = 0; u < MV_UNROLL; ++u) {
        const double x[4] = {(double)v[u].x - x0[0], (double)v[u].y - x0[1], (double)v[u].z - x0[2],
                             (double)v[u].w - x0[3]};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          a1[u & 1][k] += x[k];
          a2[u & 1][k] = fma(x[k], x[k], a2[u & 1][k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double s1 = a1[0][k] + a1[1][k], s2 = a2[0][k] + a2[1][k];
      const double mean = x0[k] + s1 / n;
      const double var = n_draws > 1 ? (s2 - s1 * s1 / n) / (n - 1.0) : 0.0;
      s_mean[k] += mean;
      s_mean2[k] = fma(mean, mean, s_mean2[k]);
      s_var[k] += var;
    }
  }
  const int64_t D = 4 * d4, d = 4 * q;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    atomicAdd(sums + d + k, s_mean[k]);
    atomicAdd(sums + D + d + k, s_mean2[k]);
    atomicAdd(sums + 2 * D + d + k, s_var[k]);
  }
}

// Lags shift + 1 .. shift + VW in ONE pass: a thread owns 2 dimensions and keeps VW consecutive
// draws of its chain in a register ring (as doubles, converted once); draw i is paired with the
// VW draws that precede draw i - shift.  For shift == 0 (lags 1 .. VW) every trace element is read
// once; later lag blocks read a second, delayed stream of the same column (shift draws behind)
// that feeds the ring.  Either way an element costs one subtraction and one FMA per lag: for
// many lags the kernel is bound by the FP64 pipe (2 * VW instructions per element), not by HBM.
// (The per-lag kernel this replaces for lag0 > 1 re-read the whole trace for every lag: 83 lags
// of the 17 GB shard took 470 ms; as 6 blocks of 16 lags they take ~70 ms.)
constexpr int VW = 16;
constexpr int VW_THREADS = 128;
constexpr int VW_CHAINS = 64;

// v: the current draws (index base + k), w: the draws that enter the ring (v itself, or the
// delayed stream)
template <bool FIRST, bool DELAYED>
__device__ __forceinline__ void variogram_block(const float2 (&v)[VW], const float2 (&w)[VW], int64_t base,
                                                int64_t n_draws, double (&ring)[VW][2], double (&acc)[VW][2]) {
#pragma unroll
  for (int k = 0; k < VW; ++k) {
    if (base + k < n_draws) {
      const double x[2] = {(double)v[k].x, (double)v[k].y};
#pragma unroll
      for (int t = 1; t <= VW; ++t) {
        if (!FIRST || k >= t) {                  // (FIRST: draw k has no predecessor at lag t > k)
          const int r = (k - t) & (VW - 1);
          const double d0 = x[0] - ring[r][0], d1 = x[1] - ring[r][1];
          acc[t - 1][0] = fma(d0, d0, acc[t - 1][0]);
          acc[t - 1][1] = fma(d1, d1, acc[t - 1][1]);
        }
      }
      ring[k][0] = DELAYED ? (double)w[k].x : x[0];
      ring[k][1] = DELAYED ? (double)w[k].y : x[1];
    }
  }
}

template <bool DELAYED>
__global__ void __launch_bounds__(VW_THREADS, 2)
variogram_window_kernel(const float2* __restrict__ trace2, double* __restrict__ out, int64_t n_draws,
                        int64_t n_chains, int64_t d2, int n_lags, int64_t shift) {
  const int64_t q = (int64_t)blockIdx.x * VW_THREADS + threadIdx.x;
  if (q >= d2) return;
  const int64_t c0 = (int64_t)blockIdx.y * VW_CHAINS;
  const int64_t c1 = min(c0 + VW_CHAINS, n_chains);
  const int64_t stride = n_chains * d2;
  double acc[VW][2];
#pragma unroll
  for (int t = 0; t < VW; ++t) acc[t][0] = acc[t][1] = 0.0;
  for (int64_t j = c0; j < c1; ++j) {
    const float2* p = trace2 + j * d2 + q;
    double ring[VW][2];
    // draw i = shift + base + k is paired with draws base + k - t (t = 1 .. VW), i.e. lag shift + t
    for (int64_t base = 0; shift + base < n_draws; base += VW) {
      float2 v[VW];
#pragma unroll
      for (int k = 0; k < VW; ++k)
        v[k] = shift + base + k < n_draws ? __ldcs(p + (shift + base + k) * stride) : make_float2(0.0f, 0.0f);
      if constexpr (DELAYED) {
        float2 w[VW];
#pragma unroll
        for (int k = 0; k < VW; ++k)        // (base + k < shift + base + k: in range whenever it is used)
          w[k] = shift + base + k < n_draws ? __ldcs(p + (base + k) * stride) : make_float2(0.0f, 0.0f);
        if (base == 0) variogram_block<true, true>(v, w, shift + base, n_draws, ring, acc);
        else variogram_block<false, true>(v, w, shift + base, n_draws, ring, acc);
      } else {
        if (base == 0) variogram_block<true, false>(v, v, base, n_draws, ring, acc);
        else variogram_block<false, false>(v, v, base, n_draws, ring, acc);
      }
    }
  }
  const int64_t D = 2 * d2, d = 2 * q;
#pragma unroll
  for (int t = 0; t < VW; ++t)
    if (t < n_lags) {
      atomicAdd(out + (int64_t)t * D + d, acc[t][0]);
      atomicAdd(out + (int64_t)t * D + d + 1, acc[t][1]);
    }
}


// Variogram sums of a FEW selected dimensions (the ones whose ESS stopping rule has not fired
// after the first block of lags): thread = chain, CTA = (selected dimension, 256 chains, lag).
constexpr int VS_THREADS = 256;
__global__ void __launch_bounds__(VS_THREADS)
variogram_select_kernel(const float* __restrict__ trace, const int64_t* __restrict__ dims, double* __restrict__ out,
                        int64_t n_draws, int64_t n_chains, int64_t n_dims, int64_t n_sel, int64_t lag0) {
  __shared__ double red[VS_THREADS / 32];
  const int64_t d = dims[blockIdx.x];
  const int64_t j = (int64_t)blockIdx.y * VS_THREADS + threadIdx.x;
  const int64_t t = lag0 + blockIdx.z;
  double acc = 0.0;
  if (j < n_chains && t < n_draws) {
    const int64_t stride = n_chains * n_dims;
    const float* p = trace + j * n_dims + d;
    double a[2] = {0.0, 0.0};
    int64_t i = t;
    for (; i + 1 < n_draws; i += 2) {
      const double d0 = (double)p[i * stride] - (double)p[(i - t) * stride];
      const double d1 = (double)p[(i + 1) * stride] - (double)p[(i + 1 - t) * stride];
      a[0] = fma(d0, d0, a[0]);
      a[1] = fma(d1, d1, a[1]);
    }
    if (i < n_draws) {
      const double d0 = (double)p[i * stride] - (double)p[(i - t) * stride];
      a[0] = fma(d0, d0, a[0]);
    }
    acc = a[0] + a[1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < VS_THREADS / 32; ++w) s += red[w];
    atomicAdd(out + (int64_t)blockIdx.z * n_sel + blockIdx.x, s);
  }
}

}  // namespace sgmcmc

using namespace sgmcmc;

static int check_trace_args(const float* trace, const double* out, int64_t n_draws, int64_t n_chains,
                            int64_t n_dims) {
  SG_REQUIRE(n_draws >= 0 && n_chains >= 0 && n_dims >= 0, SGMCMC_E_INVALID, "negative size");
  SG_REQUIRE(trace && out, SGMCMC_E_INVALID, "trace and output must not be NULL");
  SG_REQUIRE((n_chains + CHAINS_PER_BLOCK - 1) / CHAINS_PER_BLOCK <= 65535, SGMCMC_E_INVALID, "too many chains");
  return SGMCMC_OK;
}

// `sums` ([3, D]) is ACCUMULATED into: zero it before the first call.
extern "C" int sgmcmc_chain_moments_f32(const float* trace, double* sums, int64_t n_draws, int64_t n_chains,
                                        int64_t n_dims, void* stream) {
  if (int rc = check_trace_args(trace, sums, n_draws, n_chains, n_dims)) return rc;
  if (n_draws == 0 || n_chains == 0 || n_dims == 0) return SGMCMC_OK;
  if (n_dims % 4 == 0 && aligned_to(trace, 16) && (n_chains + MV_CHAINS - 1) / MV_CHAINS <= 65535) {
    const int64_t d4 = n_dims / 4;
    const dim3 grid((unsigned)((d4 + MV_THREADS - 1) / MV_THREADS), (unsigned)((n_chains + MV_CHAINS - 1) / MV_CHAINS));
    chain_moments_vec_kernel<<<grid, MV_THREADS, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(trace), sums, n_draws, n_chains, d4);
    return check_launch("chain_moments_vec_kernel");
  }
  const dim3 block(MT_DX, MT_DY);
  const dim3 grid((unsigned)((n_dims + MT_DX - 1) / MT_DX),
                  (unsigned)((n_chains + CHAINS_PER_BLOCK - 1) / CHAINS_PER_BLOCK));
  chain_moments_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(trace, sums, n_draws, n_chains, n_dims);
  return check_launch("chain_moments_kernel");
}

// `variogram` ([n_lags, D]) is ACCUMULATED into: zero it before the first call.
extern "C" int sgmcmc_variogram_f32(const float* trace, double* variogram, int64_t n_draws, int64_t n_chains,
                                    int64_t n_dims, int64_t lag0, int64_t n_lags, void* stream) {
  if (int rc = check_trace_args(trace, variogram, n_draws, n_chains, n_dims)) return rc;
  SG_REQUIRE(lag0 >= 1 && n_lags >= 0 && n_lags <= 65535, SGMCMC_E_INVALID, "lag0 must be >= 1, n_lags in [0, 65535]");
  if (n_draws == 0 || n_chains == 0 || n_dims == 0 || n_lags == 0) return SGMCMC_OK;
  if (n_dims % 2 == 0 && aligned_to(trace, 8) && (n_chains + VW_CHAINS - 1) / VW_CHAINS <= 65535) {
    // blocks of VW lags, one (lags 1 .. VW: what well-mixed chains need) or two streams over the
    // trace per block
    const int64_t d2 = n_dims / 2;
    const dim3 grid((unsigned)((d2 + VW_THREADS - 1) / VW_THREADS), (unsigned)((n_chains + VW_CHAINS - 1) / VW_CHAINS));
    for (int64_t b = 0; b < n_lags; b += VW) {
      const int64_t shift = lag0 - 1 + b;
      if (shift >= n_draws) break;                       // (no pairs at these lags: the sums stay 0)
      const int nl = (int)(n_lags - b < VW ? n_lags - b : VW);
      const float2* t2 = reinterpret_cast<const float2*>(trace);
      if (shift == 0)
        variogram_window_kernel<false><<<grid, VW_THREADS, 0, (cudaStream_t)stream>>>(t2, variogram + b * n_dims,
                                                                                      n_draws, n_chains, d2, nl, 0);
      else
        variogram_window_kernel<true><<<grid, VW_THREADS, 0, (cudaStream_t)stream>>>(t2, variogram + b * n_dims,
                                                                                     n_draws, n_chains, d2, nl, shift);
      if (int rc = check_launch("variogram_window_kernel")) return rc;
    }
    return SGMCMC_OK;
  }
  const dim3 block(MT_DX, MT_DY);
  const dim3 grid((unsigned)((n_dims + MT_DX - 1) / MT_DX),
                  (unsigned)((n_chains + CHAINS_PER_BLOCK - 1) / CHAINS_PER_BLOCK), (unsigned)n_lags);
  variogram_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(trace, variogram, n_draws, n_chains, n_dims, lag0);
  return check_launch("variogram_kernel");
}

// Like sgmcmc_variogram_f32 for the `n_sel` dimensions listed in `dims` (device, int64) only:
// `variogram` is [n_lags, n_sel], ACCUMULATED into.
extern "C" int sgmcmc_variogram_select_f32(const float* trace, const int64_t* dims, double* variogram,
                                           int64_t n_draws, int64_t n_chains, int64_t n_dims, int64_t n_sel,
                                           int64_t lag0, int64_t n_lags, void* stream) {
  if (int rc = check_trace_args(trace, variogram, n_draws, n_chains, n_dims)) return rc;
  SG_REQUIRE(lag0 >= 1 && n_lags >= 0 && n_lags <= 65535 && n_sel >= 0, SGMCMC_E_INVALID,
             "lag0 must be >= 1, n_lags in [0, 65535], n_sel >= 0");
  SG_REQUIRE((n_chains + VS_THREADS - 1) / VS_THREADS <= 65535, SGMCMC_E_INVALID, "too many chains");
  if (n_draws == 0 || n_chains == 0 || n_dims == 0 || n_lags == 0 || n_sel == 0) return SGMCMC_OK;
  SG_REQUIRE(dims != nullptr, SGMCMC_E_INVALID, "dims must not be NULL");
  const dim3 grid((unsigned)n_sel, (unsigned)((n_chains + VS_THREADS - 1) / VS_THREADS), (unsigned)n_lags);
  variogram_select_kernel<<<grid, VS_THREADS, 0, (cudaStream_t)stream>>>(trace, dims, variogram, n_draws, n_chains,
                                                                         n_dims, n_sel, lag0);
  return check_launch("variogram_select_kernel");
}
