// K1-K3: streaming element-wise sampler updates over the flat [C chains x D params]
// state (HBM-bandwidth bound: 44 / 24 B per element-step for SGHMC burn-in / sampling,
// 36 / 16 for SGLD, 20 for relativistic SGHMC).
//
// One thread owns groups of 4 consecutive elements (one 128-bit access per array per
// group; one Philox4x32-10 call yields the group's 4 normals).  UNROLL groups per thread
// are loaded up-front so every thread keeps 6*UNROLL independent 16-byte loads in flight.
// State arrays are read once and written once per step, so loads/stores use the
// streaming (evict-first) cache policy.
#include "sampler_math.cuh"

namespace sgmcmc {

template <typename T>
struct Pack {
  T v[4];
};

template <typename T, bool ALIGNED>
__device__ __forceinline__ Pack<T> load_pack(const T* __restrict__ p, int64_t group, int valid) {
  Pack<T> r;
  if (ALIGNED && valid == 4) {
    if constexpr (sizeof(T) == 4) {
      const float4 q = ld_stream(reinterpret_cast<const float4*>(p) + group);
      r.v[0] = q.x; r.v[1] = q.y; r.v[2] = q.z; r.v[3] = q.w;
    } else {
      const double2 a = ld_stream(reinterpret_cast<const double2*>(p) + 2 * group);
      const double2 b = ld_stream(reinterpret_cast<const double2*>(p) + 2 * group + 1);
      r.v[0] = a.x; r.v[1] = a.y; r.v[2] = b.x; r.v[3] = b.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) r.v[i] = (i < valid) ? p[4 * group + i] : (T)1;
  }
  return r;
}

template <typename T, bool ALIGNED>
__device__ __forceinline__ void store_pack(T* __restrict__ p, int64_t group, int valid, const Pack<T>& r,
                                           bool keep = false) {
  if (ALIGNED && valid == 4) {
    if constexpr (sizeof(T) == 4) {
      // keep: write-back with normal L2 priority (the next kernel reads this array first)
      if (keep) reinterpret_cast<float4*>(p)[group] = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
      else st_stream(reinterpret_cast<float4*>(p) + group, make_float4(r.v[0], r.v[1], r.v[2], r.v[3]));
    } else {
      st_stream(reinterpret_cast<double2*>(p) + 2 * group, make_double2(r.v[0], r.v[1]));
      st_stream(reinterpret_cast<double2*>(p) + 2 * group + 1, make_double2(r.v[2], r.v[3]));
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < valid) p[4 * group + i] = r.v[i];
  }
}

template <typename T, bool EXT_Z, bool ALIGNED>
__device__ __forceinline__ Pack<T> noise_pack(const T* __restrict__ z, int64_t group, int valid,
                                              uint64_t group_offset, uint64_t step, uint64_t seed) {
  if constexpr (EXT_Z) {
    return load_pack<T, ALIGNED>(z, group, valid);
  } else {
    float f[4];
    normal4((uint64_t)group + group_offset, step, seed, f);
    Pack<T> r;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.v[i] = (T)f[i];
    return r;
  }
}

// ------------------------------------------------------------------------------------
// K1  SGHMC
// ------------------------------------------------------------------------------------
template <typename T, bool BURN_IN, bool STORE_MINV, bool EXT_Z, bool ALIGNED, int UNROLL>
__global__ void sghmc_update_kernel(T* __restrict__ theta, T* __restrict__ v, T* __restrict__ tau,
                                    T* __restrict__ g, T* __restrict__ v_hat, T* __restrict__ minv,
                                    const T* __restrict__ grad, const T* __restrict__ z, int64_t n,
                                    SghmcScalars<T> s, NoiseArgs na) {
  const int64_t n_groups = (n + 3) >> 2;
  // grid-stride over the groups: with the default launch the loop runs once; a capped
  // (persistent) grid walks the array in chunks of gridDim.x * blockDim.x * UNROLL groups
  const int64_t chunk = (int64_t)gridDim.x * blockDim.x * UNROLL;
  // na.reverse: the CTAs scheduled first take the END of the array -- what the kernel that ran
  // before (K4, ascending over chains) touched last and is therefore still in L2
  const unsigned bid = (na.reverse & 1) ? gridDim.x - 1u - blockIdx.x : blockIdx.x;
  for (int64_t base = (int64_t)bid * ((int64_t)blockDim.x * UNROLL) + threadIdx.x;
       base - threadIdx.x < n_groups; base += chunk) {
  Pack<T> th[UNROLL], vv[UNROLL], gr[UNROLL], ta[UNROLL], gg[UNROLL], vh[UNROLL], mi[UNROLL], zz[UNROLL];
  int valid[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    const int64_t gi = base + (int64_t)u * blockDim.x;
    const int64_t rem = n - 4 * gi;
    valid[u] = gi < n_groups ? (rem >= 4 ? 4 : (int)rem) : 0;
    if (valid[u] > 0) {
      th[u] = load_pack<T, ALIGNED>(theta, gi, valid[u]);
      vv[u] = load_pack<T, ALIGNED>(v, gi, valid[u]);
      gr[u] = load_pack<T, ALIGNED>(grad, gi, valid[u]);
      if constexpr (BURN_IN) {
        ta[u] = load_pack<T, ALIGNED>(tau, gi, valid[u]);
        gg[u] = load_pack<T, ALIGNED>(g, gi, valid[u]);
        vh[u] = load_pack<T, ALIGNED>(v_hat, gi, valid[u]);
      } else {
        mi[u] = load_pack<T, ALIGNED>(minv, gi, valid[u]);
      }
      zz[u] = noise_pack<T, EXT_Z, ALIGNED>(z, gi, valid[u], na.group_offset, na.step, na.seed);
    }
  }
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    if (valid[u] > 0) {
      const int64_t gi = base + (int64_t)u * blockDim.x;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        T minv_t;
        if constexpr (BURN_IN) {
          minv_t = adapt(ta[u].v[i], gg[u].v[i], vh[u].v[i], gr[u].v[i]);
          mi[u].v[i] = minv_t;
        } else {
          minv_t = mi[u].v[i];
        }
        sghmc_apply(th[u].v[i], vv[u].v[i], minv_t, gr[u].v[i], zz[u].v[i], s);
      }
      store_pack<T, ALIGNED>(theta, gi, valid[u], th[u], (na.reverse & 2) != 0);
      store_pack<T, ALIGNED>(v, gi, valid[u], vv[u]);
      if constexpr (BURN_IN) {
        store_pack<T, ALIGNED>(tau, gi, valid[u], ta[u]);
        store_pack<T, ALIGNED>(g, gi, valid[u], gg[u]);
        store_pack<T, ALIGNED>(v_hat, gi, valid[u], vh[u]);
        if constexpr (STORE_MINV) store_pack<T, ALIGNED>(minv, gi, valid[u], mi[u]);
      }
    }
  }
  }
}

// ------------------------------------------------------------------------------------
// K2  SGLD
// ------------------------------------------------------------------------------------
template <typename T, bool BURN_IN, bool STORE_MINV, bool EXT_Z, bool ALIGNED, int UNROLL>
__global__ void sgld_update_kernel(T* __restrict__ theta, T* __restrict__ tau, T* __restrict__ g,
                                   T* __restrict__ v_hat, T* __restrict__ minv,
                                   const T* __restrict__ grad, const T* __restrict__ z, int64_t n,
                                   SgldScalars<T> s, NoiseArgs na) {
  const int64_t n_groups = (n + 3) >> 2;
  // grid-stride over the groups: with the default launch the loop runs once; a capped
  // (persistent) grid walks the array in chunks of gridDim.x * blockDim.x * UNROLL groups
  const int64_t chunk = (int64_t)gridDim.x * blockDim.x * UNROLL;
  for (int64_t base = (int64_t)blockIdx.x * ((int64_t)blockDim.x * UNROLL) + threadIdx.x;
       base - threadIdx.x < n_groups; base += chunk) {
  Pack<T> th[UNROLL], gr[UNROLL], ta[UNROLL], gg[UNROLL], vh[UNROLL], mi[UNROLL], zz[UNROLL];
  int valid[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    const int64_t gi = base + (int64_t)u * blockDim.x;
    const int64_t rem = n - 4 * gi;
    valid[u] = gi < n_groups ? (rem >= 4 ? 4 : (int)rem) : 0;
    if (valid[u] > 0) {
      th[u] = load_pack<T, ALIGNED>(theta, gi, valid[u]);
      gr[u] = load_pack<T, ALIGNED>(grad, gi, valid[u]);
      if constexpr (BURN_IN) {
        ta[u] = load_pack<T, ALIGNED>(tau, gi, valid[u]);
        gg[u] = load_pack<T, ALIGNED>(g, gi, valid[u]);
        vh[u] = load_pack<T, ALIGNED>(v_hat, gi, valid[u]);
      } else {
        mi[u] = load_pack<T, ALIGNED>(minv, gi, valid[u]);
      }
      zz[u] = noise_pack<T, EXT_Z, ALIGNED>(z, gi, valid[u], na.group_offset, na.step, na.seed);
    }
  }
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    if (valid[u] > 0) {
      const int64_t gi = base + (int64_t)u * blockDim.x;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        T minv_t;
        if constexpr (BURN_IN) {
          minv_t = adapt(ta[u].v[i], gg[u].v[i], vh[u].v[i], gr[u].v[i]);
          mi[u].v[i] = minv_t;
        } else {
          minv_t = mi[u].v[i];
        }
        sgld_apply(th[u].v[i], minv_t, gr[u].v[i], zz[u].v[i], s);
      }
      store_pack<T, ALIGNED>(theta, gi, valid[u], th[u]);
      if constexpr (BURN_IN) {
        store_pack<T, ALIGNED>(tau, gi, valid[u], ta[u]);
        store_pack<T, ALIGNED>(g, gi, valid[u], gg[u]);
        store_pack<T, ALIGNED>(v_hat, gi, valid[u], vh[u]);
        if constexpr (STORE_MINV) store_pack<T, ALIGNED>(minv, gi, valid[u], mi[u]);
      }
    }
  }
  }
}

// ------------------------------------------------------------------------------------
// K3  relativistic SGHMC
// ------------------------------------------------------------------------------------
template <typename T, bool UNIT, bool EXT_Z, bool ALIGNED, int UNROLL>
__global__ void rsghmc_update_kernel(T* __restrict__ theta, T* __restrict__ p,
                                     const T* __restrict__ grad, const T* __restrict__ z, int64_t n,
                                     RsghmcScalars<T> s, NoiseArgs na) {
  const int64_t n_groups = (n + 3) >> 2;
  // grid-stride over the groups: with the default launch the loop runs once; a capped
  // (persistent) grid walks the array in chunks of gridDim.x * blockDim.x * UNROLL groups
  const int64_t chunk = (int64_t)gridDim.x * blockDim.x * UNROLL;
  for (int64_t base = (int64_t)blockIdx.x * ((int64_t)blockDim.x * UNROLL) + threadIdx.x;
       base - threadIdx.x < n_groups; base += chunk) {
  Pack<T> th[UNROLL], pp[UNROLL], gr[UNROLL], zz[UNROLL];
  int valid[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    const int64_t gi = base + (int64_t)u * blockDim.x;
    const int64_t rem = n - 4 * gi;
    valid[u] = gi < n_groups ? (rem >= 4 ? 4 : (int)rem) : 0;
    if (valid[u] > 0) {
      th[u] = load_pack<T, ALIGNED>(theta, gi, valid[u]);
      pp[u] = load_pack<T, ALIGNED>(p, gi, valid[u]);
      gr[u] = load_pack<T, ALIGNED>(grad, gi, valid[u]);
      zz[u] = noise_pack<T, EXT_Z, ALIGNED>(z, gi, valid[u], na.group_offset, na.step, na.seed);
    }
  }
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    if (valid[u] > 0) {
      const int64_t gi = base + (int64_t)u * blockDim.x;
#pragma unroll
      for (int i = 0; i < 4; ++i) rsghmc_apply<T, UNIT>(th[u].v[i], pp[u].v[i], gr[u].v[i], zz[u].v[i], s);
      store_pack<T, ALIGNED>(theta, gi, valid[u], th[u]);
      store_pack<T, ALIGNED>(p, gi, valid[u], pp[u]);
    }
  }
  }
}

template <bool ALIGNED>
__global__ void normal_fill_kernel(float* __restrict__ out, int64_t n, NoiseArgs na) {
  const int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gi >= ((n + 3) >> 2)) return;
  const int64_t rem = n - 4 * gi;
  const int valid = rem >= 4 ? 4 : (int)rem;
  const Pack<float> r = noise_pack<float, false, ALIGNED>(nullptr, gi, valid, na.group_offset, na.step, na.seed);
  store_pack<float, ALIGNED>(out, gi, valid, r);
}

// ------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------
struct LaunchShape {
  int threads, unroll;
  unsigned blocks;
};

template <typename T>
static inline LaunchShape launch_shape(int64_t n) {
  LaunchShape ls;
  ls.threads = tuning_threads();
  ls.unroll = tuning_unroll();
  // the double kernels are only instantiated at unroll 1 (see SG_DISPATCH_UNROLL) and
  // need up to ~170 registers per thread
  if (sizeof(T) == 8) { ls.unroll = 1; if (ls.threads > 256) ls.threads = 256; }
  const int64_t n_groups = (n + 3) / 4;
  const int64_t per_block = (int64_t)ls.threads * ls.unroll;
  ls.blocks = (unsigned)((n_groups + per_block - 1) / per_block);
  const int cap = tuning_update_max_ctas();
  if (cap > 0 && ls.blocks > (unsigned)cap) ls.blocks = (unsigned)cap;
  return ls;
}

#define SG_DISPATCH_UNROLL(UN, ...)                                     \
  if constexpr (sizeof(T) == 8) {                                       \
    constexpr int U = 1; __VA_ARGS__;                                   \
  } else {                                                              \
    switch (UN) {                                                       \
      case 2: { constexpr int U = 2; __VA_ARGS__; } break;              \
      default: { constexpr int U = 1; __VA_ARGS__; } break;             \
    }                                                                   \
  }

#define SG_DISPATCH_BOOL(VAL, NAME, ...) \
  if (VAL) { constexpr bool NAME = true; __VA_ARGS__; } else { constexpr bool NAME = false; __VA_ARGS__; }

static int check_common(int64_t n, uint64_t elem_offset) {
  SG_REQUIRE(n >= 0, SGMCMC_E_INVALID, "n must be >= 0 (got %lld)", (long long)n);
  SG_REQUIRE(elem_offset % 4 == 0, SGMCMC_E_INVALID, "elem_offset must be a multiple of 4");
  SG_REQUIRE((n + 3) / 4 / 128 < 0x7fffffffLL, SGMCMC_E_INVALID, "n too large for one launch");
  return SGMCMC_OK;
}

template <typename T>
static int sghmc_step(T* theta, T* v, T* tau, T* g, T* v_hat, T* minv, const T* grad, const T* z,
                      int64_t n, T epsilon, T mdecay, T scale_grad, int burn_in, int store_minv,
                      uint64_t seed, uint64_t step, uint64_t elem_offset, void* stream) {
  if (int rc = check_common(n, elem_offset)) return rc;
  if (n == 0) return SGMCMC_OK;
  SG_REQUIRE(theta && v && grad, SGMCMC_E_INVALID, "sghmc_step: theta, v and grad must not be NULL");
  SG_REQUIRE(!burn_in || (tau && g && v_hat), SGMCMC_E_INVALID,
             "sghmc_step: burn-in needs tau, g and v_hat");
  SG_REQUIRE((burn_in && !store_minv) || minv, SGMCMC_E_INVALID, "sghmc_step: minv must not be NULL");
  SG_REQUIRE(scale_grad > 0, SGMCMC_E_INVALID, "sghmc_step: scale_grad must be > 0");
  const void* ptrs[] = {theta, v, tau, g, v_hat, minv, grad, z};
  bool aligned = true;
  for (const void* p : ptrs) {
    SG_REQUIRE(aligned_to(p, sizeof(T)), SGMCMC_E_ALIGN, "sghmc_step: pointer not aligned to element size");
    aligned = aligned && aligned_to(p, 16);
  }
  const SghmcScalars<T> s = make_sghmc_scalars<T>(epsilon, mdecay, scale_grad);
  NoiseArgs na{seed, step, elem_offset / 4};
  na.reverse = tuning_update_reverse();
  const LaunchShape ls = launch_shape<T>(n);
  cudaStream_t st = (cudaStream_t)stream;
  const bool ext_z = z != nullptr;
  const int carve = tuning_update_carveout();   // 100: share SMs with K4 (two-stream pipeline); -1: default
#define SG_LAUNCH_K1(BI, SM)                                                                          \
  {                                                                                                   \
    auto k = sghmc_update_kernel<T, BI, SM, EZ, AL, U>;                                               \
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, carve);                   \
    k<<<ls.blocks, ls.threads, 0, st>>>(theta, v, tau, g, v_hat, minv, grad, z, n, s, na);            \
  }
  SG_DISPATCH_UNROLL(ls.unroll,
    SG_DISPATCH_BOOL(aligned, AL,
      SG_DISPATCH_BOOL(ext_z, EZ,
        if (burn_in) {
          if (store_minv) SG_LAUNCH_K1(true, true)
          else SG_LAUNCH_K1(true, false)
        } else {
          SG_LAUNCH_K1(false, false)
        })))
#undef SG_LAUNCH_K1
  return check_launch("sghmc_update_kernel");
}

template <typename T>
static int sgld_step(T* theta, T* tau, T* g, T* v_hat, T* minv, const T* grad, const T* z, int64_t n,
                     T epsilon, T A, T scale_grad, int burn_in, int store_minv, uint64_t seed,
                     uint64_t step, uint64_t elem_offset, void* stream) {
  if (int rc = check_common(n, elem_offset)) return rc;
  if (n == 0) return SGMCMC_OK;
  SG_REQUIRE(theta && grad, SGMCMC_E_INVALID, "sgld_step: theta and grad must not be NULL");
  SG_REQUIRE(!burn_in || (tau && g && v_hat), SGMCMC_E_INVALID, "sgld_step: burn-in needs tau, g and v_hat");
  SG_REQUIRE((burn_in && !store_minv) || minv, SGMCMC_E_INVALID, "sgld_step: minv must not be NULL");
  const void* ptrs[] = {theta, tau, g, v_hat, minv, grad, z};
  bool aligned = true;
  for (const void* p : ptrs) {
    SG_REQUIRE(aligned_to(p, sizeof(T)), SGMCMC_E_ALIGN, "sgld_step: pointer not aligned to element size");
    aligned = aligned && aligned_to(p, 16);
  }
  const SgldScalars<T> s = make_sgld_scalars<T>(epsilon, A, scale_grad);
  const NoiseArgs na{seed, step, elem_offset / 4};
  const LaunchShape ls = launch_shape<T>(n);
  cudaStream_t st = (cudaStream_t)stream;
  const bool ext_z = z != nullptr;
  SG_DISPATCH_UNROLL(ls.unroll,
    SG_DISPATCH_BOOL(aligned, AL,
      SG_DISPATCH_BOOL(ext_z, EZ,
        if (burn_in) {
          if (store_minv)
            sgld_update_kernel<T, true, true, EZ, AL, U><<<ls.blocks, ls.threads, 0, st>>>(
                theta, tau, g, v_hat, minv, grad, z, n, s, na);
          else
            sgld_update_kernel<T, true, false, EZ, AL, U><<<ls.blocks, ls.threads, 0, st>>>(
                theta, tau, g, v_hat, minv, grad, z, n, s, na);
        } else {
          sgld_update_kernel<T, false, false, EZ, AL, U><<<ls.blocks, ls.threads, 0, st>>>(
              theta, tau, g, v_hat, minv, grad, z, n, s, na);
        })))
  return check_launch("sgld_update_kernel");
}

template <typename T>
static int rsghmc_step(T* theta, T* p, const T* grad, const T* z, int64_t n, T epsilon, T mass, T c,
                       T D, T Bhat, uint64_t seed, uint64_t step, uint64_t elem_offset, void* stream) {
  if (int rc = check_common(n, elem_offset)) return rc;
  if (n == 0) return SGMCMC_OK;
  SG_REQUIRE(theta && p && grad, SGMCMC_E_INVALID, "rsghmc_step: theta, p and grad must not be NULL");
  const void* ptrs[] = {theta, p, grad, z};
  bool aligned = true;
  for (const void* q : ptrs) {
    SG_REQUIRE(aligned_to(q, sizeof(T)), SGMCMC_E_ALIGN, "rsghmc_step: pointer not aligned to element size");
    aligned = aligned && aligned_to(q, 16);
  }
  const RsghmcScalars<T> s = make_rsghmc_scalars<T>(epsilon, mass, c, D, Bhat);
  const NoiseArgs na{seed, step, elem_offset / 4};
  const LaunchShape ls = launch_shape<T>(n);
  cudaStream_t st = (cudaStream_t)stream;
  const bool ext_z = z != nullptr;
  const bool unit = mass == (T)1 && c == (T)1;
  SG_DISPATCH_UNROLL(ls.unroll,
    SG_DISPATCH_BOOL(aligned, AL,
      SG_DISPATCH_BOOL(ext_z, EZ,
        SG_DISPATCH_BOOL(unit, UN1,
          rsghmc_update_kernel<T, UN1, EZ, AL, U><<<ls.blocks, ls.threads, 0, st>>>(theta, p, grad, z, n, s, na)))))
  return check_launch("rsghmc_update_kernel");
}

}  // namespace sgmcmc

using namespace sgmcmc;

extern "C" {

int sgmcmc_sghmc_step_f32(float* theta, float* v, float* tau, float* g, float* v_hat, float* minv,
                          const float* grad, const float* z, int64_t n, float epsilon, float mdecay,
                          float scale_grad, int burn_in, int store_minv, uint64_t seed, uint64_t step,
                          uint64_t elem_offset, void* stream) {
  return sghmc_step<float>(theta, v, tau, g, v_hat, minv, grad, z, n, epsilon, mdecay, scale_grad,
                           burn_in, store_minv, seed, step, elem_offset, stream);
}
int sgmcmc_sghmc_step_f64(double* theta, double* v, double* tau, double* g, double* v_hat, double* minv,
                          const double* grad, const double* z, int64_t n, double epsilon, double mdecay,
                          double scale_grad, int burn_in, int store_minv, uint64_t seed, uint64_t step,
                          uint64_t elem_offset, void* stream) {
  return sghmc_step<double>(theta, v, tau, g, v_hat, minv, grad, z, n, epsilon, mdecay, scale_grad,
                            burn_in, store_minv, seed, step, elem_offset, stream);
}
int sgmcmc_sgld_step_f32(float* theta, float* tau, float* g, float* v_hat, float* minv,
                         const float* grad, const float* z, int64_t n, float epsilon, float A,
                         float scale_grad, int burn_in, int store_minv, uint64_t seed, uint64_t step,
                         uint64_t elem_offset, void* stream) {
  return sgld_step<float>(theta, tau, g, v_hat, minv, grad, z, n, epsilon, A, scale_grad, burn_in,
                          store_minv, seed, step, elem_offset, stream);
}
int sgmcmc_sgld_step_f64(double* theta, double* tau, double* g, double* v_hat, double* minv,
                         const double* grad, const double* z, int64_t n, double epsilon, double A,
                         double scale_grad, int burn_in, int store_minv, uint64_t seed, uint64_t step,
                         uint64_t elem_offset, void* stream) {
  return sgld_step<double>(theta, tau, g, v_hat, minv, grad, z, n, epsilon, A, scale_grad, burn_in,
                           store_minv, seed, step, elem_offset, stream);
}
int sgmcmc_rsghmc_step_f32(float* theta, float* p, const float* grad_cost, const float* z, int64_t n,
                           float epsilon, float mass, float speed_of_light, float D, float Bhat,
                           uint64_t seed, uint64_t step, uint64_t elem_offset, void* stream) {
  return rsghmc_step<float>(theta, p, grad_cost, z, n, epsilon, mass, speed_of_light, D, Bhat, seed,
                            step, elem_offset, stream);
}
int sgmcmc_rsghmc_step_f64(double* theta, double* p, const double* grad_cost, const double* z, int64_t n,
                           double epsilon, double mass, double speed_of_light, double D, double Bhat,
                           uint64_t seed, uint64_t step, uint64_t elem_offset, void* stream) {
  return rsghmc_step<double>(theta, p, grad_cost, z, n, epsilon, mass, speed_of_light, D, Bhat, seed,
                             step, elem_offset, stream);
}

int sgmcmc_normal_fill_f32(float* out, int64_t n, uint64_t seed, uint64_t step, uint64_t elem_offset,
                           void* stream) {
  if (int rc = check_common(n, elem_offset)) return rc;
  if (n == 0) return SGMCMC_OK;
  SG_REQUIRE(out, SGMCMC_E_INVALID, "normal_fill: out must not be NULL");
  SG_REQUIRE(aligned_to(out, 4), SGMCMC_E_ALIGN, "normal_fill: pointer not aligned");
  const NoiseArgs na{seed, step, elem_offset / 4};
  const int threads = 256;
  const unsigned blocks = (unsigned)(((n + 3) / 4 + threads - 1) / threads);
  if (aligned_to(out, 16))
    normal_fill_kernel<true><<<blocks, threads, 0, (cudaStream_t)stream>>>(out, n, na);
  else
    normal_fill_kernel<false><<<blocks, threads, 0, (cudaStream_t)stream>>>(out, n, na);
  return check_launch("normal_fill_kernel");
}

}  // extern "C"
