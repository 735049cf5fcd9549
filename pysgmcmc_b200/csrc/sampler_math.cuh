// Per-element sampler arithmetic, shared by the streaming update kernels (K1-K3),
// the fused target-chain kernel (K6) and the fused BNN-SGHMC kernel (K5).
// Every line mirrors one TensorFlow op of the reference, in the reference's
// parenthesisation, with un-contracted IEEE arithmetic (see common.cuh).
#pragma once

#include "common.cuh"

namespace sgmcmc {

// Philox counter of a launch: element group (4 elements) `i` of the launch's arrays draws
// from counter (i + group_offset, step) under key `seed`.
struct NoiseArgs {
  uint64_t seed, step, group_offset;
  int reverse = 0;   // K1: walk the array from its end (L2 reuse after K4, see bnn.cu)
};

// ---- host-side scalar prefixes -------------------------------------------------------
// Computed in T with the reference's operation order (oracle/samplers.py:sghmc_scalars);
// built with -ffp-contract=off so the host compiler does not fuse them either.
template <typename T>
struct SghmcScalars {
  T a;         // (2 * eps_s^2) * mdecay          sghmc.py:211-214
  T c4;        // eps_s^4                         sghmc.py:216
  T neg_eps2;  // -(eps^2), UNSCALED eps          sghmc.py:235
  T mdecay;
};

template <typename T>
inline T host_pow(T x, T y);
template <>
inline float host_pow<float>(float x, float y) { return powf(x, y); }
template <>
inline double host_pow<double>(double x, double y) { return pow(x, y); }
template <typename T>
inline T host_sqrt(T x);
template <>
inline float host_sqrt<float>(float x) { return sqrtf(x); }
template <>
inline double host_sqrt<double>(double x) { return sqrt(x); }

template <typename T>
inline SghmcScalars<T> make_sghmc_scalars(T epsilon, T mdecay, T scale_grad) {
  SghmcScalars<T> s;
  volatile T eps_s = epsilon / host_sqrt<T>(scale_grad);          // sghmc.py:115
  volatile T e2 = host_pow<T>(eps_s, (T)2);
  volatile T two_e2 = (T)2 * e2;
  s.a = two_e2 * mdecay;
  s.c4 = host_pow<T>(eps_s, (T)4);
  s.neg_eps2 = -host_pow<T>(epsilon, (T)2);
  s.mdecay = mdecay;
  return s;
}

template <typename T>
struct SgldScalars {
  T two_eps;      // 2 * eps                      sgld.py:187
  T A;            // (A - noise), noise == 0      sgld.py:189
  T scale_den;    // scale_grad + (2*sign*c + c)  tensor_utils.py:269
  T neg_eps;      // -eps                         sgld.py:203
};

template <typename T>
inline SgldScalars<T> make_sgld_scalars(T epsilon, T A, T scale_grad) {
  SgldScalars<T> s;
  const T c = (T)1e-16;
  volatile T two_c = (T)2 * (T)((scale_grad > 0) - (scale_grad < 0)) * c;
  volatile T off = two_c + c;
  s.two_eps = (T)2 * epsilon;
  volatile T a_minus_noise = A - (T)0;
  s.A = a_minus_noise;
  s.scale_den = scale_grad + off;
  s.neg_eps = -epsilon;
  return s;
}

template <typename T>
struct RsghmcScalars {
  T eps, m, m2c2, D, noise_sigma;   // noise_sigma = sqrt(eps * (2*D - eps*Bhat))
};

template <typename T>
inline RsghmcScalars<T> make_rsghmc_scalars(T epsilon, T mass, T c, T D, T Bhat) {
  RsghmcScalars<T> s;
  volatile T m2 = mass * mass, c2 = c * c;
  s.eps = epsilon;
  s.m = mass;
  s.m2c2 = m2 * c2;                                   // square(m) * square(c)   :123
  s.D = D;
  volatile T two_d = (T)2 * D;
  volatile T eb = epsilon * Bhat;
  volatile T inner = two_d - eb;
  volatile T prod = epsilon * inner;
  s.noise_sigma = host_sqrt<T>(prod);                 // :125
  return s;
}

// ---- burn-in adaptation: sghmc.py:165-196 == sgld.py:153-180 ---------------------------
// In: OLD tau, g, v_hat.  Out: minv_t and the NEW tau, g, v_hat.
template <typename T>
__device__ __forceinline__ T adapt(T& tau, T& g, T& v_hat, T grad) {
  using F = ieee<T>;
  const T one = (T)1;
  const T r_t = F::div(one, F::add(tau, one));                                    // :168
  const T tau_t = F::add(tau, F::add(safe_divide(F::mul(F::mul(-g, g), tau), v_hat), one));  // :172-176
  const T minv_t = safe_divide(one, safe_sqrt(v_hat));                            // :179-183
  const T g_t = F::add(g, F::add(F::mul(-r_t, g), F::mul(r_t, grad)));            // :186-190
  const T v_hat_t = F::add(v_hat, F::add(F::mul(-r_t, v_hat), F::mul(r_t, F::mul(grad, grad))));  // :192-196
  tau = tau_t; g = g_t; v_hat = v_hat_t;
  return minv_t;
}

// ---- SGHMC: sghmc.py:211-243 -----------------------------------------------------------
// (the `- 2 eps_s^3 minv^2 * noise` term is exactly 0 for finite minv and is omitted)
template <typename T>
__device__ __forceinline__ void sghmc_apply(T& theta, T& v, T minv_t, T grad, T z,
                                            const SghmcScalars<T>& s) {
  using F = ieee<T>;
  const T noise_scale = F::sub(F::mul(s.a, minv_t), s.c4);                        // :211-217
  const T sigma = F::sqrt(F::max(noise_scale, (T)1e-16));                         // :220
  const T sample = F::mul(sigma, z);                                              // base_classes.py:218
  const T drift = F::sub(F::mul(F::mul(s.neg_eps2, minv_t), grad), F::mul(s.mdecay, v));
  const T v_t = F::add(v, F::add(drift, sample));                                 // :233-238
  v = v_t;
  theta = F::add(theta, v_t);                                                     // :241-243
}

// ---- SGLD: sgld.py:186-204 ---------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void sgld_apply(T& theta, T minv_t, T grad, T z, const SgldScalars<T>& s) {
  using F = ieee<T>;
  const T sigma = safe_sqrt(F::mul(s.two_eps, F::div(F::mul(minv_t, s.A), s.scale_den)));   // :186-191
  const T sample = F::mul(sigma, z);
  const T drift = F::mul(F::mul(F::mul(s.neg_eps, minv_t), s.A), grad);                     // :203
  theta = F::add(theta, F::add(drift, sample));                                             // :201-204
}

// ---- relativistic SGHMC: relativistic_sghmc.py:120-135 ----------------------------------
// UNIT: mass == 1 and speed_of_light == 1 (the reference's defaults).  x / 1 and 1 * x are
// exact in IEEE arithmetic, so skipping them is bit-identical and saves two of the four
// divisions per element.
template <typename T, bool UNIT>
__device__ __forceinline__ T rel_velocity(T p, const RsghmcScalars<T>& s) {
  using F = ieee<T>;
  // eps * p / (m * sqrt(p*p / (m^2 c^2) + 1))
  if constexpr (UNIT) {
    return F::div(F::mul(s.eps, p), F::sqrt(F::add(F::mul(p, p), (T)1)));
  } else {
    return F::div(F::mul(s.eps, p),
                  F::mul(s.m, F::sqrt(F::add(F::div(F::mul(p, p), s.m2c2), (T)1))));
  }
}
template <typename T, bool UNIT = false>
__device__ __forceinline__ void rsghmc_apply(T& theta, T& p, T grad_cost, T z,
                                             const RsghmcScalars<T>& s) {
  using F = ieee<T>;
  const T grad = -grad_cost;                                                       // :100-103
  const T p_grad = rel_velocity<T, UNIT>(p, s);                                    // :123
  const T n = F::mul(s.noise_sigma, z);                                            // :125
  const T p_t = F::add(p, F::sub(F::add(F::mul(s.eps, grad), n), F::mul(s.D, p_grad)));   // :126-129
  p = p_t;
  theta = F::add(theta, rel_velocity<T, UNIT>(p_t, s));                            // :131-135
}

}  // namespace sgmcmc
