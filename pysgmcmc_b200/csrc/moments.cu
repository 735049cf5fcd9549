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
  const dim3 block(MT_DX, MT_DY);
  const dim3 grid((unsigned)((n_dims + MT_DX - 1) / MT_DX),
                  (unsigned)((n_chains + CHAINS_PER_BLOCK - 1) / CHAINS_PER_BLOCK), (unsigned)n_lags);
  variogram_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(trace, variogram, n_draws, n_chains, n_dims, lag0);
  return check_launch("variogram_kernel");
}
