// K6: whole chains on a built-in test density, many steps per launch.
//
// One thread owns one chain (D = 2 banana, D = 1 gmm): the target's gradient
// (pysgmcmc/diagnostics/objective_functions.py:49-98, differentiated by hand like
// oracle/targets.py), the sampler update (sampler_math.cuh) and the noise are fused and
// the chain state stays in registers for `n_steps` steps.  HBM traffic is the state once
// per launch plus the (optional) thinned trace, so this kernel is bound by the
// per-chain dependent-instruction latency, not by bandwidth: throughput scales with the
// number of resident chains.
#include "sampler_math.cuh"

namespace sgmcmc {

struct GmmParams {
  float lw[3];    // log w_i
  float h[3];     // -0.5 * log(2 pi var_i)
  float mu[3];
  float var[3];
};

// cost = -loglik and d cost / d theta, in oracle/targets.py's operation order
struct Banana {
  static constexpr int D = 2;
  static __device__ __forceinline__ float cost_grad(const float* x, float* grad, const GmmParams&) {
    using F = ieee<float>;
    const float x0 = x[0], x1 = x[1];
    const float u = F::sub(F::add(x1, F::mul(F::mul(0.1f, x0), x0)), 10.0f);
    const float cost = F::mul(0.5f, F::add(F::mul(F::mul(0.01f, x0), x0), F::mul(u, u)));
    grad[0] = F::add(F::mul(0.01f, x0), F::mul(F::mul(0.2f, x0), u));
    grad[1] = u;
    return cost;
  }
};

struct Gmm {
  static constexpr int D = 1;
  static __device__ __forceinline__ float cost_grad(const float* x, float* grad, const GmmParams& P) {
    using F = ieee<float>;
    float comp[3], d[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      d[i] = F::sub(x[0], P.mu[i]);
      const float q = F::div(F::mul(0.5f, F::mul(d[i], d[i])), P.var[i]);
      comp[i] = F::add(P.lw[i], F::sub(P.h[i], q));
    }
    const float m = fmaxf(fmaxf(comp[0], comp[1]), comp[2]);
    float e[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) e[i] = expf(F::sub(comp[i], m));
    const float s = F::add(F::add(e[0], e[1]), e[2]);
    const float cost = -F::add(m, logf(s));
    float gsum = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i) gsum = F::add(gsum, F::mul(F::div(e[i], s), F::div(d[i], P.var[i])));
    grad[0] = gsum;
    return cost;
  }
};

struct ChainRunArgs {
  float *theta, *a1, *tau, *g, *v_hat, *minv;
  const float* z;
  float *trace, *cost_trace;
  int64_t n_chains, n_steps, n_burn_in, keep_every;
  int adapt_forever;
  uint64_t seed, step0, chain_offset;
  SghmcScalars<float> sghmc;
  SgldScalars<float> sgld;
  RsghmcScalars<float> rsghmc;
  GmmParams gmm;
};

template <int SAMPLER, typename Target>
__global__ void __launch_bounds__(64) target_chains_kernel(ChainRunArgs a) {
  constexpr int D = Target::D;
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.n_chains) return;
  float theta[D], a1[D], tau[D], g[D], v_hat[D], minv[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    theta[d] = a.theta[c * D + d];
    if (SAMPLER != SGMCMC_SAMPLER_SGLD) a1[d] = a.a1[c * D + d];
    if (SAMPLER != SGMCMC_SAMPLER_RSGHMC) {
      tau[d] = a.tau[c * D + d];
      g[d] = a.g[c * D + d];
      v_hat[d] = a.v_hat[c * D + d];
      minv[d] = a.minv[c * D + d];
    }
  }
  const uint64_t e0 = (a.chain_offset + (uint64_t)c) * D;   // global flat index of element 0
  for (int64_t s = 0; s < a.n_steps; ++s) {
    float grad[D], zz[D];
    const float cost = Target::cost_grad(theta, grad, a.gmm);
    if (a.z != nullptr) {
#pragma unroll
      for (int d = 0; d < D; ++d) zz[d] = a.z[(s * a.n_chains + c) * D + d];
    } else {
      float f[4];
      normal4(e0 >> 2, a.step0 + (uint64_t)s, a.seed, f);
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const int comp = (int)((e0 + d) & 3);
        zz[d] = comp == 0 ? f[0] : comp == 1 ? f[1] : comp == 2 ? f[2] : f[3];
      }
    }
    const bool adaptive = a.adapt_forever || s < a.n_burn_in;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      if (SAMPLER == SGMCMC_SAMPLER_RSGHMC) {
        rsghmc_apply(theta[d], a1[d], grad[d], zz[d], a.rsghmc);
      } else {
        if (adaptive) minv[d] = adapt(tau[d], g[d], v_hat[d], grad[d]);
        if (SAMPLER == SGMCMC_SAMPLER_SGHMC) sghmc_apply(theta[d], a1[d], minv[d], grad[d], zz[d], a.sghmc);
        else sgld_apply(theta[d], minv[d], grad[d], zz[d], a.sgld);
      }
    }
    if ((s + 1) % a.keep_every == 0) {
      const int64_t k = (s + 1) / a.keep_every - 1;
      if (a.trace != nullptr) {
#pragma unroll
        for (int d = 0; d < D; ++d) a.trace[(k * a.n_chains + c) * D + d] = theta[d];
      }
      if (a.cost_trace != nullptr) a.cost_trace[k * a.n_chains + c] = cost;
    }
  }
#pragma unroll
  for (int d = 0; d < D; ++d) {
    a.theta[c * D + d] = theta[d];
    if (SAMPLER != SGMCMC_SAMPLER_SGLD) a.a1[c * D + d] = a1[d];
    if (SAMPLER != SGMCMC_SAMPLER_RSGHMC) {
      a.tau[c * D + d] = tau[d];
      a.g[c * D + d] = g[d];
      a.v_hat[c * D + d] = v_hat[d];
      a.minv[c * D + d] = minv[d];
    }
  }
}

static GmmParams make_gmm(int target) {
  // objective_functions.py:62-98; constants rounded to float like oracle/targets.py
  const double var1[3] = {1.0, 1.0, 1.0};
  const double var2[3] = {1.0 / 0.5, 0.5, 1.0 / 0.5};
  const double var3[3] = {1.0 / 0.3, 0.3, 1.0 / 0.3};
  const double* var = target == SGMCMC_TARGET_GMM1 ? var1 : target == SGMCMC_TARGET_GMM2 ? var2 : var3;
  const double mu[3] = {-5.0, 0.0, 5.0};
  GmmParams P;
  for (int i = 0; i < 3; ++i) {
    P.lw[i] = (float)log(1.0 / 3.0);
    P.h[i] = -0.5f * (float)log(2.0 * M_PI * var[i]);
    P.mu[i] = (float)mu[i];
    P.var[i] = (float)var[i];
  }
  return P;
}


// ---------------------------------------------------------------------------------------------
// K6 for SVGD: a small particle set (n <= 128) on a built-in test density, many steps per launch in
// ONE CTA.  Everything a step needs -- gradients, the n x n squared distances, the exact median
// (radix select over the shared-memory copy), the RBF kernel, the Stein direction, the AdaGrad
// history and the update (pysgmcmc/samplers/svgd.py:125-182, oracle/svgd.py) -- stays in shared
// memory and registers; HBM sees the particles once per launch plus the thinned trace.  This is
// the reference notebook's regime (docs/source/notebooks/SVGD.ipynb: 10 particles on the banana,
// 50 000 steps), which is launch-latency bound when run kernel by kernel.
// Arithmetic and summation orders follow the FFMA kernels of svgd.cu except the kernel row sums
// (sequential here, a tree there): the two paths agree to fp32 rounding, not bit for bit.
// ---------------------------------------------------------------------------------------------
struct SvgdRunArgs {
  float *X, *hist, *trace, *cost_trace;
  int n;
  int64_t n_steps, keep_every;
  float eps, alpha, one_minus_alpha, fudge;
  GmmParams gmm;
};

__device__ __forceinline__ uint32_t svgd_float_key(uint32_t b) { return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u); }
__device__ __forceinline__ uint32_t svgd_key_float(uint32_t k) { return k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu); }

template <typename Target>
__global__ void __launch_bounds__(256) svgd_target_run_kernel(SvgdRunArgs a) {
  constexpr int D = Target::D;
  extern __shared__ __align__(16) float sm[];
  const int n = a.n, tid = threadIdx.x;
  float* P = sm;                      // [n, n] distances, then the kernel matrix
  float* x = P + n * n;               // [n, D]
  float* g = x + n * D;               // [n, D]
  float* ksum = g + n * D;            // [n]
  float* cost = ksum + n;             // [n]
  __shared__ uint32_t h[2][256];
  __shared__ uint32_t prefix[2], rank[2];
  __shared__ float bw[2];             // h, h^2

  const bool owner = tid < n * D;     // one thread per coordinate
  const int oi = owner ? tid / D : 0;
  float hreg = owner ? a.hist[tid] : 0.0f;
  if (owner) x[tid] = a.X[tid];
  __syncthreads();
  const uint32_t n_values = (uint32_t)(n * n);
  const float nf = (float)n;

  for (int64_t step = 0; step < a.n_steps; ++step) {
    // gradients of the COST at the current particles (svgd.py:125)
    if (tid < n) {
      float gi[D];
      cost[tid] = Target::cost_grad(&x[tid * D], gi, a.gmm);
#pragma unroll
      for (int d = 0; d < D; ++d) g[tid * D + d] = gi[d];
    }
    // squared distances (svgd.py:151-152), difference first
    for (int idx = tid; idx < n * n; idx += 256) {
      const int i = idx / n, j = idx - i * n;
      float s = 0.0f;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float df = x[i * D + d] - x[j * D + d];
        s = fmaf(df, df, s);
      }
      const float nrm = __fsqrt_rn(s);
      P[idx] = __fmul_rn(nrm, nrm);
    }
    if (tid == 0) {
      rank[1] = n_values / 2;
      rank[0] = (n_values & 1u) ? n_values / 2 : n_values / 2 - 1;
      prefix[0] = prefix[1] = 0;
    }
    __syncthreads();
    // exact median: 4 passes of 8 bits over order-preserving keys (tensor_utils.py:194-208)
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      h[0][tid] = 0;
      h[1][tid] = 0;
      __syncthreads();
      const uint32_t p0 = prefix[0], p1 = prefix[1];
      const uint32_t mask = (shift == 24) ? 0u : (0xFFFFFFFFu << (shift + 8));
      for (int idx = tid; idx < n * n; idx += 256) {
        const uint32_t key = svgd_float_key(__float_as_uint(P[idx]));
        const uint32_t bin = (key >> shift) & 255u;
        if ((key & mask) == p0) atomicAdd(&h[0][bin], 1u);
        if (p0 != p1 && (key & mask) == p1) atomicAdd(&h[1][bin], 1u);
      }
      __syncthreads();
      if (tid < 64) {                                   // warp 0 narrows rank 0, warp 1 rank 1
        const int w = tid >> 5;
        uint32_t bin, before;
        warp_pick_bin((p0 == p1) ? h[0] : h[w], rank[w], bin, before);
        if ((tid & 31) == 0) {
          prefix[w] |= bin << shift;
          rank[w] -= before;
        }
      }
      __syncthreads();
    }
    if (tid == 0) {
      const float lo = __uint_as_float(svgd_key_float(prefix[0])), hi = __uint_as_float(svgd_key_float(prefix[1]));
      const float med = (prefix[0] == prefix[1]) ? lo : __fdiv_rn(__fadd_rn(hi, lo), 2.0f);
      const float hb = __fsqrt_rn(__fdiv_rn(__fmul_rn(0.5f, med), logf(__fadd_rn(nf, 1.0f))));   // svgd.py:155-157
      bw[0] = hb;
      bw[1] = __fmul_rn(hb, hb);
    }
    __syncthreads();
    const float h2 = bw[1];
    for (int idx = tid; idx < n * n; idx += 256) P[idx] = expf(__fdiv_rn(__fdiv_rn(-P[idx], h2), 2.0f));
    __syncthreads();
    if (tid < n) {
      float s = 0.0f;
      for (int j = 0; j < n; ++j) s += P[tid * n + j];
      ksum[tid] = s;
    }
    __syncthreads();
    // Stein direction, AdaGrad history, update (svgd.py:130-148,162-167)
    float xnew = 0.0f;
    if (owner) {
      const int d = tid - oi * D;
      float kg = 0.0f, kx = 0.0f;
      for (int j = 0; j < n; ++j) {
        const float k = P[oi * n + j];
        kg = fmaf(k, g[j * D + d], kg);
        kx = fmaf(k, x[j * D + d], kx);
      }
      const float xv = x[tid];
      const float kgrad = __fdiv_rn(__fadd_rn(-kx, __fmul_rn(xv, ksum[oi])), h2);
      const float phi = __fdiv_rn(__fadd_rn(kg, kgrad), nf);
      hreg = __fadd_rn(__fmul_rn(a.alpha, hreg), __fmul_rn(a.one_minus_alpha, __fmul_rn(phi, phi)));
      const float adj = __fdiv_rn(phi, __fadd_rn(a.fudge, __fsqrt_rn(hreg)));
      xnew = __fsub_rn(xv, __fmul_rn(a.eps, adj));
    }
    __syncthreads();                                    // every thread has read the old particles
    if (owner) x[tid] = xnew;
    if ((step + 1) % a.keep_every == 0) {
      const int64_t k = (step + 1) / a.keep_every - 1;
      if (owner && a.trace) a.trace[k * n * D + tid] = xnew;
      if (tid < n && a.cost_trace) a.cost_trace[k * n + tid] = cost[tid];   // cost of the PRE-update particles
    }
    __syncthreads();
  }
  if (owner) {
    a.X[tid] = x[tid];
    a.hist[tid] = hreg;
  }
}

}  // namespace sgmcmc

using namespace sgmcmc;

extern "C" int sgmcmc_target_chains_run_f32(int sampler, int target, float* theta, float* a1, float* tau,
                                            float* g, float* v_hat, float* minv, const float* z,
                                            float* trace, float* cost_trace, int64_t n_chains,
                                            int64_t n_steps, int64_t n_burn_in, int adapt_forever,
                                            int64_t keep_every, const sgmcmc_hyper_t* hyper,
                                            uint64_t seed, uint64_t step0, uint64_t chain_offset,
                                            void* stream) {
  SG_REQUIRE(sampler >= 0 && sampler <= 2, SGMCMC_E_INVALID, "unknown sampler id %d", sampler);
  SG_REQUIRE(target >= 0 && target <= 3, SGMCMC_E_INVALID, "unknown target id %d", target);
  SG_REQUIRE(n_chains >= 0 && n_steps >= 0 && n_burn_in >= 0, SGMCMC_E_INVALID, "negative size");
  SG_REQUIRE(keep_every >= 1, SGMCMC_E_INVALID, "keep_every must be >= 1");
  SG_REQUIRE(hyper != nullptr && theta != nullptr, SGMCMC_E_INVALID, "hyper/theta must not be NULL");
  SG_REQUIRE(sampler == SGMCMC_SAMPLER_SGLD || a1 != nullptr, SGMCMC_E_INVALID, "a1 (V or p) must not be NULL");
  SG_REQUIRE(sampler == SGMCMC_SAMPLER_RSGHMC || (tau && g && v_hat && minv), SGMCMC_E_INVALID,
             "tau, g, v_hat and minv must not be NULL for burn-in samplers");
  const int D = target == SGMCMC_TARGET_BANANA ? 2 : 1;
  SG_REQUIRE((chain_offset * D) % 4 == 0, SGMCMC_E_INVALID, "chain_offset*D must be a multiple of 4");
  if (n_chains == 0 || n_steps == 0) return SGMCMC_OK;
  ChainRunArgs a;
  a.theta = theta; a.a1 = a1; a.tau = tau; a.g = g; a.v_hat = v_hat; a.minv = minv;
  a.z = z; a.trace = trace; a.cost_trace = cost_trace;
  a.n_chains = n_chains; a.n_steps = n_steps; a.n_burn_in = n_burn_in; a.keep_every = keep_every;
  a.adapt_forever = adapt_forever; a.seed = seed; a.step0 = step0; a.chain_offset = chain_offset;
  a.sghmc = make_sghmc_scalars<float>(hyper->epsilon, hyper->mdecay, hyper->scale_grad);
  a.sgld = make_sgld_scalars<float>(hyper->epsilon, hyper->A, hyper->scale_grad);
  a.rsghmc = make_rsghmc_scalars<float>(hyper->epsilon, hyper->mass, hyper->speed_of_light, hyper->D, hyper->Bhat);
  a.gmm = make_gmm(target);
  const int threads = 64;
  const unsigned blocks = (unsigned)((n_chains + threads - 1) / threads);
  cudaStream_t st = (cudaStream_t)stream;
#define SG_LAUNCH_TC(S)                                                                  \
  if (target == SGMCMC_TARGET_BANANA) target_chains_kernel<S, Banana><<<blocks, threads, 0, st>>>(a); \
  else target_chains_kernel<S, Gmm><<<blocks, threads, 0, st>>>(a);
  if (sampler == SGMCMC_SAMPLER_SGHMC) { SG_LAUNCH_TC(SGMCMC_SAMPLER_SGHMC) }
  else if (sampler == SGMCMC_SAMPLER_SGLD) { SG_LAUNCH_TC(SGMCMC_SAMPLER_SGLD) }
  else { SG_LAUNCH_TC(SGMCMC_SAMPLER_RSGHMC) }
#undef SG_LAUNCH_TC
  return check_launch("target_chains_kernel");
}

extern "C" int sgmcmc_svgd_target_run_f32(int target, float* particles, float* historical_grad, float* trace,
                                          float* cost_trace, int64_t n_particles, int64_t n_steps,
                                          int64_t keep_every, float epsilon, float alpha, float one_minus_alpha,
                                          float fudge_factor, void* stream) {
  SG_REQUIRE(target >= 0 && target <= 3, SGMCMC_E_INVALID, "unknown target id %d", target);
  SG_REQUIRE(n_particles >= 1 && n_particles <= 128, SGMCMC_E_UNSUPPORTED,
             "the fused SVGD kernel holds at most 128 particles (got %lld)", (long long)n_particles);
  SG_REQUIRE(n_steps >= 0 && keep_every >= 1, SGMCMC_E_INVALID, "n_steps must be >= 0 and keep_every >= 1");
  SG_REQUIRE(particles && historical_grad, SGMCMC_E_INVALID, "particles / historical_grad must not be NULL");
  if (n_steps == 0) return SGMCMC_OK;
  const int n = (int)n_particles, D = target == SGMCMC_TARGET_BANANA ? 2 : 1;
  SvgdRunArgs a;
  a.X = particles; a.hist = historical_grad; a.trace = trace; a.cost_trace = cost_trace;
  a.n = n; a.n_steps = n_steps; a.keep_every = keep_every;
  a.eps = epsilon; a.alpha = alpha; a.one_minus_alpha = one_minus_alpha; a.fudge = fudge_factor;
  a.gmm = make_gmm(target);
  const size_t smem = sizeof(float) * ((size_t)n * n + 2 * (size_t)n * D + 2 * (size_t)n);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
  if (target == SGMCMC_TARGET_BANANA) {
    e = cudaFuncSetAttribute(svgd_target_run_kernel<Banana>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    if (e != cudaSuccess) return set_error(SGMCMC_E_CUDA, "svgd_target_run_kernel: %s", cudaGetErrorString(e));
    svgd_target_run_kernel<Banana><<<1, 256, smem, st>>>(a);
  } else {
    e = cudaFuncSetAttribute(svgd_target_run_kernel<Gmm>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    if (e != cudaSuccess) return set_error(SGMCMC_E_CUDA, "svgd_target_run_kernel: %s", cudaGetErrorString(e));
    svgd_target_run_kernel<Gmm><<<1, 256, smem, st>>>(a);
  }
  return check_launch("svgd_target_run_kernel");
}
