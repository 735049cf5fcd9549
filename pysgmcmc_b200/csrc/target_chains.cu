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
