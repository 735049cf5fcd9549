// Shared device/host helpers for libsgmcmc_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sgmcmc_b200.h"

namespace sgmcmc {

// ---- error plumbing (capi.cu) ------------------------------------------------------
int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);
void count_launch();
int tuning_threads();
int tuning_unroll();
int bnn_variant_count();
void set_bnn_variant(int v);
void set_bnn_max_ctas(int n);
void set_bnn_chunk(int64_t c);
void set_bnn_pipeline(int64_t chunk, int ring);
int tuning_update_carveout();
int tuning_update_max_ctas();
int tuning_update_reverse();

#define SG_REQUIRE(cond, code, ...)                          \
  do {                                                       \
    if (!(cond)) return ::sgmcmc::set_error(code, __VA_ARGS__); \
  } while (0)

inline bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// ---- IEEE round-to-nearest arithmetic that the compiler may NOT contract -----------
// TensorFlow evaluates the reference's update op by op (no FMA), so parity with the
// float32 oracle needs every product and sum rounded on its own.  The intrinsics
// below are never fused, independent of -fmad.
template <typename T>
struct ieee;

template <>
struct ieee<float> {
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
  static __device__ __forceinline__ float max(float a, float b) { return fmaxf(a, b); }
  static __device__ __forceinline__ float min(float a, float b) { return fminf(a, b); }
};

template <>
struct ieee<double> {
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
  static __device__ __forceinline__ double max(double a, double b) { return fmax(a, b); }
  static __device__ __forceinline__ double min(double a, double b) { return fmin(a, b); }
};

// safe_divide(x, y) = x / (y + (2*sign(y)*c + c)), c = 1e-16  (tensor_utils.py:269).
// (2*sign*c + c) evaluates to fl(3c) for y>0, c for y==0, -c for y<0.
template <typename T>
__device__ __forceinline__ T safe_den(T y) {
  const T c = (T)1e-16;
  const T c3 = ieee<T>::add(ieee<T>::mul((T)2.0, c), c);
  const T off = (y > (T)0) ? c3 : ((y < (T)0) ? -c : c);
  return ieee<T>::add(y, off);   // NaN y stays NaN
}
template <typename T>
__device__ __forceinline__ T safe_divide(T x, T y) {
  return ieee<T>::div(x, safe_den(y));
}
// safe_sqrt(x) = sqrt(clip(x, 0, inf))  (tensor_utils.py:319-323)
template <typename T>
__device__ __forceinline__ T safe_sqrt(T x) {
  return ieee<T>::sqrt(ieee<T>::max(x, (T)0));
}

// ---- Philox4x32-10 (Salmon et al. 2011) + Box-Muller: the engine's noise stream ----
// Restated for the tests in oracle/philox.py (pinned there by Random123 vectors).
struct Philox4 {
  uint32_t x, y, z, w;
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ float u32_to_unit_open(uint32_t x) {
  // float32(x) * 2^-32 + 2^-33 in (0, 1]; the product is exact, one rounding in the add
  return __fadd_rn(__fmul_rn(__uint2float_rn(x), 2.3283064365386963e-10f), 1.1641532182693481e-10f);
}

// Four N(0,1) draws for elements 4*group .. 4*group+3 of step `step`.
// Box-Muller on the SFU: lg2.approx / sqrt.approx / sin.approx / cos.approx (MUFU), about
// 12 instructions per pair.  Absolute error of a draw ~1e-6, worst seen 8e-6 (|sin|,|cos| error 2^-21.4
// times a radius <= 6.8), far below the statistical resolution of any chain; the tests
// compare against oracle/philox.py with atol 2e-5 (worst seen 8e-6).  The uniforms are bit-exact.
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void normal4(uint64_t group, uint64_t step, uint64_t seed, float out[4]) {
  const Philox4 r = philox4x32_10((uint32_t)group, (uint32_t)(group >> 32), (uint32_t)step,
                                  (uint32_t)(step >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
  const float u0 = u32_to_unit_open(r.x), u1 = u32_to_unit_open(r.y);
  const float u2 = u32_to_unit_open(r.z), u3 = u32_to_unit_open(r.w);
  // -2 ln u = (-2 ln 2) * log2 u
  const float r0 = fast_sqrt(-1.3862943611198906f * __log2f(u0));
  const float r1 = fast_sqrt(-1.3862943611198906f * __log2f(u2));
  const float a0 = 6.283185307179586f * u1, a1 = 6.283185307179586f * u3;
  out[0] = r0 * __cosf(a0); out[1] = r0 * __sinf(a0);
  out[2] = r1 * __cosf(a1); out[3] = r1 * __sinf(a1);
}

// ---- radix-select helper: which of 256 histogram bins holds the element of rank `r`? -----------
// Executed by one full warp (lane l scans bins 8l .. 8l+7, a shuffle scan finds the crossing lane).
// Returns the bin (255 if r is beyond the total) and the number of elements in the bins before it.
__device__ __forceinline__ void warp_pick_bin(const uint32_t* hist, uint32_t r, uint32_t& bin, uint32_t& before) {
  const int lane = (int)(threadIdx.x & 31);
  uint32_t c[8], s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    c[k] = hist[lane * 8 + k];
    s += c[k];
  }
  uint32_t incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= o) incl += v;
  }
  const unsigned crossing = __ballot_sync(0xFFFFFFFFu, r < incl);
  const int L = crossing ? __ffs(crossing) - 1 : 31;
  uint32_t b = 7, cum = incl - s;
  if (lane == L) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (r < cum + c[k]) { b = (uint32_t)k; break; }
      cum += c[k];
    }
  }
  bin = __shfl_sync(0xFFFFFFFFu, (uint32_t)L * 8u + b, L);
  before = __shfl_sync(0xFFFFFFFFu, cum, L);
}

// ---- streaming global memory access (single-use data: do not keep it in L1) --------
template <typename V>
__device__ __forceinline__ V ld_stream(const V* p) { return __ldcs(p); }
template <typename V>
__device__ __forceinline__ void st_stream(V* p, const V& v) { __stcs(p, v); }

}  // namespace sgmcmc
