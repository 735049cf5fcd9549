// K5 as ONE kernel: a whole `sample, cost = next(sampler)` of BNN-SGHMC for one chain per
// CTA -- the cost + gradient of pysgmcmc/models/bayesian_neural_network.py:28-69,337-388 on
// the tensor pipe (bnn_mma.cuh) followed, in the same CTA, by the SGHMC update of
// pysgmcmc/samplers/sghmc.py:165-251 (sampler_math.cuh, the arithmetic of K1 op for op).
//
// Why fuse.  As two kernels the step moves 52 B per element through HBM (K4: theta in,
// gradient out; K1: 44 B) and the two kernels cannot overlap: K4 is issue / tensor bound with
// HBM idle, K1 is HBM bound with the tensor pipe idle.  Here the gradient never leaves shared
// memory (it is produced in place over the staged parameters) and theta is re-read from L2
// moments after it was staged, so the step costs 40 B per element of HBM traffic (20 after
// burn-in), and the 6 CTAs resident on an SM are in different phases: while some wait on
// their state arrays, others run their MMAs.
//
// The update phase of a chain is D / 4 = 1313 groups of 4 elements walked by the CTA's
// threads, software-pipelined in registers (two stages of UPD_U groups per thread), with the
// chain's state rows prefetched into L2 by TMA while the MMAs run; one Philox4x32-10 call per
// group -- the SAME (group, step) -> counter mapping as K1, so the
// fused step and K4 + K1 produce bit-identical states (tests/test_bnn_gpu.py).
#include "bnn_mma.cuh"

namespace sgmcmc {

constexpr int UPD_U = 2;     // groups per thread per pipeline stage (two stages in flight)

__device__ __forceinline__ void unpack4(const float4& q, float (&r)[4]) {
  r[0] = q.x; r[1] = q.y; r[2] = q.z; r[3] = q.w;
}
__device__ __forceinline__ float4 pack4(const float (&r)[4]) { return make_float4(r[0], r[1], r[2], r[3]); }

// One element group's state: theta, V and (burn-in) tau, g, v_hat or (sampling) the frozen minv in `a`.
template <bool BURN_IN>
struct GroupState;
template <>
struct GroupState<true> { float4 th, v, a, g, h; };
template <>
struct GroupState<false> { float4 th, v, a; };

// TMA bulk prefetch of `bytes` (a multiple of 16) at a 16-byte aligned address into L2.
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Start moving the chain's state rows towards L2 while its MMAs run (one thread).
template <bool BURN_IN>
__device__ __forceinline__ void prefetch_chain_state(const FusedStepArgs& f, int64_t chain, int D) {
  const uint32_t bytes = (uint32_t)D * 4u;
  const int64_t o = chain * D;
  prefetch_l2_bulk(f.v + o, bytes);
  if constexpr (BURN_IN) {
    prefetch_l2_bulk(f.tau + o, bytes);
    prefetch_l2_bulk(f.g + o, bytes);
    prefetch_l2_bulk(f.v_hat + o, bytes);
  } else {
    prefetch_l2_bulk(f.minv + o, bytes);
  }
}

// The SGHMC update of chain `chain` with its gradient in shared memory (R, parameter layout).
// Two register stages of UPD_U groups per thread: the loads of stage i+1 are in flight
// while stage i runs its ~150 instructions per element, so a warp rarely waits on memory
// even with only 12 warps on the SM.
template <bool BURN_IN, int NTHR>
__device__ __forceinline__ void sghmc_update_chain(const FusedStepArgs& f, const float* __restrict__ R,
                                                   int64_t chain, int D, int tid) {
  const int n4 = D >> 2;
  const int64_t g0 = chain * n4;                       // first element group of this chain
  float4* th4 = reinterpret_cast<float4*>(f.theta) + g0;
  float4* v4 = reinterpret_cast<float4*>(f.v) + g0;
  float4* ta4 = reinterpret_cast<float4*>(f.tau) + g0;
  float4* gg4 = reinterpret_cast<float4*>(f.g) + g0;
  float4* vh4 = reinterpret_cast<float4*>(f.v_hat) + g0;
  float4* mi4 = reinterpret_cast<float4*>(f.minv) + g0;
  const float4* z4 = f.z != nullptr ? reinterpret_cast<const float4*>(f.z) + g0 : nullptr;
  const float4* gr4 = reinterpret_cast<const float4*>(R);
  const SghmcScalars<float> s = f.s;
  const bool store_minv = f.store_minv != 0;
  using Stage = GroupState<BURN_IN>[UPD_U];

  auto load = [&](Stage& b, int q0) {
#pragma unroll
    for (int u = 0; u < UPD_U; ++u) {
      const int q = q0 + u * NTHR + tid;
      if (q < n4) {
        b[u].th = __ldcg(th4 + q);                     // staged moments ago by this CTA: an L2 hit
        b[u].v = ld_stream(v4 + q);
        if constexpr (BURN_IN) {
          b[u].a = ld_stream(ta4 + q);
          b[u].g = ld_stream(gg4 + q);
          b[u].h = ld_stream(vh4 + q);
        } else {
          b[u].a = ld_stream(mi4 + q);                 // the frozen mass matrix (base_classes.py:448-454)
        }
      }
    }
  };
  auto update = [&](Stage& b, int q0) {
#pragma unroll
    for (int u = 0; u < UPD_U; ++u) {
      const int q = q0 + u * NTHR + tid;
      if (q < n4) {
        float zf[4], t[4], v[4], a[4], g[4], h[4], gr[4], mi[4];
        if (z4 != nullptr) unpack4(ld_stream(z4 + q), zf);
        else normal4((uint64_t)(g0 + q) + f.na.group_offset, f.na.step, f.na.seed, zf);
        unpack4(gr4[q], gr);
        unpack4(b[u].th, t); unpack4(b[u].v, v); unpack4(b[u].a, a);
        if constexpr (BURN_IN) { unpack4(b[u].g, g); unpack4(b[u].h, h); }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float minv_t;
          if constexpr (BURN_IN) {
            minv_t = adapt(a[i], g[i], h[i], gr[i]);
            mi[i] = minv_t;
          } else {
            minv_t = a[i];
          }
          sghmc_apply(t[i], v[i], minv_t, gr[i], zf[i], s);
        }
        st_stream(th4 + q, pack4(t));
        st_stream(v4 + q, pack4(v));
        if constexpr (BURN_IN) {
          st_stream(ta4 + q, pack4(a));
          st_stream(gg4 + q, pack4(g));
          st_stream(vh4 + q, pack4(h));
          if (store_minv) st_stream(mi4 + q, pack4(mi));
        }
      }
    }
  };

  constexpr int STEP = UPD_U * NTHR;
  GroupState<BURN_IN> b0[UPD_U], b1[UPD_U];
  load(b0, 0);
#pragma unroll 1
  for (int q0 = 0; q0 < n4; q0 += 2 * STEP) {
    load(b1, q0 + STEP);
    update(b0, q0);
    load(b0, q0 + 2 * STEP);
    update(b1, q0 + STEP);
  }
}

template <int NB8, bool BURN_IN>
__global__ void __launch_bounds__(32 * ((NB8 + 1) / 2), NB8 > 2 ? 6 : 8)
bnn_sghmc_fused_kernel(BnnArgs a, FusedStepArgs f) {
  constexpr int NW = (NB8 + 1) / 2;
  constexpr int NTHR = 32 * NW;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int D = a.L.D;
  const BnnMmaSmem s = bnn_mma_carve(smem, a.batch, a.L.n_in, D);
  for (int64_t chain = blockIdx.x; chain < a.n_chains; chain += gridDim.x) {
    float cost = 0.0f, sse = 0.0f;
    if (f.prefetch && tid == 0) prefetch_chain_state<BURN_IN>(f, chain, D);
    bnn_chain_mma<NB8, true, true, MMA_DEFAULT_MODE>(a, f.theta + chain * D, a.starts != nullptr ? a.starts + chain : nullptr,
                                   s, cost, sse);
    if (tid == 0) {
      a.cost[chain] = cost;                            // U(theta_{t-1}): base_classes.py:298-300
      if (a.mse != nullptr) a.mse[chain] = sse / (float)a.batch;
    }
    sghmc_update_chain<BURN_IN, NTHR>(f, s.R, chain, D, tid);
    chain_barrier<NW>();                               // R is restaged by the next chain
  }
}

// ---- the warp-specialised form of the same step -----------------------------------------------------
// One persistent CTA of 8 warps, two per SM: warps 0-3 are two MMA groups (2 warps each, one chain at a
// time per group: bnn_chain_mma, unchanged), warps 4-7 are update warps that apply K1's arithmetic to the
// chains the groups have finished.  A group owns two gradient buffers in shared memory: while the update
// warps consume chain j from one, the group computes chain j + 1 into the other, so the issue-bound
// gradient and the HBM-bound update overlap INSIDE an SM at all times instead of in alternating kernels
// (or in a CTA that does one after the other, above), and the gradient never touches HBM (40 B per
// element and step).  Hand-over through named barriers (bar.arrive / bar.sync, 64 + 128 threads):
//   ready[g][b]: group g -> update warps (the gradient of its current chain is complete in buffer b)
//   free [g][b]: update warps -> group g (buffer b may be overwritten)
// 128 registers per thread (K4 runs as fast at 128 as at 168: profiles/r02_k4_occupancy_experiment.jsonl),
// 110 KB of shared memory per CTA.  Bit-identical to K4 then K1 (same noise counters, same arithmetic).
constexpr int WS_GROUP_THREADS = 64, WS_UPDATE_THREADS = 128, WS_THREADS = 2 * WS_GROUP_THREADS + WS_UPDATE_THREADS;

__device__ __forceinline__ void named_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__host__ __device__ inline int bnn_ws_group_floats(int batch, int n_in, int D) {
  return ((D + 3) & ~3) + bnn_mma_smem_floats(batch, n_in, D);      // second gradient buffer + one chain's K4 buffers
}

template <int NB8, bool BURN_IN>
__global__ void __launch_bounds__(WS_THREADS, 2) bnn_sghmc_ws_kernel(BnnArgs a, FusedStepArgs f) {
  static_assert((NB8 + 1) / 2 == 2, "two MMA warps per chain");
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int D = a.L.D, D4 = (D + 3) & ~3;
  const int group_floats = bnn_ws_group_floats(a.batch, a.L.n_in, D);
  const int64_t stride = (int64_t)gridDim.x * 2;
  constexpr int HANDOVER = WS_GROUP_THREADS + WS_UPDATE_THREADS;
  if (warp < 4) {
    // ------------------------------------------------------------------ MMA group g
    const int g = warp >> 1, tid0 = g * WS_GROUP_THREADS;
    float* base = smem + g * group_floats;                          // [R0 | R1 | P | Q | Zb | X | y | scratch]
    BnnMmaSmem s = bnn_mma_carve(base + D4, a.batch, a.L.n_in, D);
    int j = 0;
    for (int64_t chain = (int64_t)blockIdx.x * 2 + g; chain < a.n_chains; chain += stride, ++j) {
      const int b = j & 1;
      if (j >= 2) named_sync(7 + 2 * g + b, HANDOVER);              // the update of chain j - 2 has read buffer b
      s.R = base + b * D4;
      if (f.prefetch && tid == tid0) prefetch_chain_state<BURN_IN>(f, chain, D);
      float cost = 0.0f, sse = 0.0f;
      bnn_chain_mma<NB8, true, true, MMA_DEFAULT_MODE>(
          a, f.theta + chain * D, a.starts != nullptr ? a.starts + chain : nullptr, s, cost, sse, tid0, 1 + g);
      if (tid == tid0) {
        a.cost[chain] = cost;                                        // U(theta_{t-1}): base_classes.py:298-300
        if (a.mse != nullptr) a.mse[chain] = sse / (float)a.batch;
      }
      named_arrive(3 + 2 * g + b, HANDOVER);                         // gradient of `chain` complete in buffer b
    }
  } else {
    // ------------------------------------------------------------------ update warps
    const int ut = tid - 2 * WS_GROUP_THREADS;
    for (int j = 0;; ++j) {
      const int b = j & 1;
      bool any = false;
#pragma unroll 1
      for (int g = 0; g < 2; ++g) {
        const int64_t chain = (int64_t)blockIdx.x * 2 + g + (int64_t)j * stride;
        if (chain >= a.n_chains) continue;
        any = true;
        named_sync(3 + 2 * g + b, HANDOVER);
        sghmc_update_chain<BURN_IN, WS_UPDATE_THREADS>(f, smem + g * group_floats + b * D4, chain, D, ut);
        if (chain + 2 * stride < a.n_chains) named_arrive(7 + 2 * g + b, HANDOVER);
      }
      if (!any) break;
    }
  }
}

static int g_bnn_fused = 0;   // 0: K4 then K1; 1: one CTA does both in turn; 2: warp-specialised (see DESIGN.md "K5")
static int g_fused_max_ctas = 0;
static int g_fused_prefetch = 1;
void set_bnn_fused_max_ctas(int n) { g_fused_max_ctas = n; }
void set_bnn_fused_prefetch(int on) { g_fused_prefetch = on; }
int bnn_fused_enabled() { return g_bnn_fused; }
void set_bnn_fused(int on) { g_bnn_fused = on; }

bool bnn_fused_supported(const BnnArgs& a, const FusedStepArgs& f) {
  if (a.batch > 32 || (a.L.D & 3) != 0) return false;
  const void* ptrs[] = {f.theta, f.v, f.tau, f.g, f.v_hat, f.minv, f.z};
  for (const void* p : ptrs)
    if (!aligned_to(p, 16)) return false;              // (NULL z is aligned)
  return (size_t)bnn_mma_smem_floats(a.batch, a.L.n_in, a.L.D) * sizeof(float) <= 227 * 1024;
}

template <int NB8>
static int launch_ws(const BnnArgs& a, const FusedStepArgs& f, cudaStream_t st) {
  const size_t smem = (size_t)2 * bnn_ws_group_floats(a.batch, a.L.n_in, a.L.D) * sizeof(float);
  SG_REQUIRE(smem <= 113 * 1024, SGMCMC_E_UNSUPPORTED, "warp-specialised BNN-SGHMC step: %zu B of shared memory per CTA", smem);
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
  }
  unsigned blocks = (unsigned)((a.n_chains + 1) / 2);
  const unsigned cap = g_fused_max_ctas > 0 ? (unsigned)g_fused_max_ctas : 2u * (unsigned)n_sm;   // persistent: 2 CTAs per SM
  if (blocks > cap) blocks = cap;
  if (f.burn_in) {
    auto k = bnn_sghmc_ws_kernel<NB8, true>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    k<<<blocks, WS_THREADS, smem, st>>>(a, f);
  } else {
    auto k = bnn_sghmc_ws_kernel<NB8, false>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    k<<<blocks, WS_THREADS, smem, st>>>(a, f);
  }
  return check_launch("bnn_sghmc_ws_kernel");
}

template <int NB8>
static int launch_fused(const BnnArgs& a, const FusedStepArgs& f, cudaStream_t st) {
  constexpr int NTHR = 32 * ((NB8 + 1) / 2);
  const size_t smem = (size_t)bnn_mma_smem_floats(a.batch, a.L.n_in, a.L.D) * sizeof(float);
  unsigned blocks = (unsigned)a.n_chains;
  if (g_fused_max_ctas > 0 && blocks > (unsigned)g_fused_max_ctas) blocks = (unsigned)g_fused_max_ctas;
  if (f.burn_in) {
    auto k = bnn_sghmc_fused_kernel<NB8, true>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    k<<<blocks, NTHR, smem, st>>>(a, f);
  } else {
    auto k = bnn_sghmc_fused_kernel<NB8, false>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    k<<<blocks, NTHR, smem, st>>>(a, f);
  }
  return check_launch("bnn_sghmc_fused_kernel");
}

int launch_bnn_sghmc_fused(const BnnArgs& a, const FusedStepArgs& f_in, cudaStream_t st) {
  FusedStepArgs f = f_in;
  f.prefetch = g_fused_prefetch;
  SG_REQUIRE(bnn_fused_supported(a, f), SGMCMC_E_UNSUPPORTED,
             "fused BNN-SGHMC step: needs batch <= 32, D %% 4 == 0 and 16-byte aligned state");
  if (g_bnn_fused == 2 && a.batch > 16) {
    const size_t smem_ws = (size_t)2 * bnn_ws_group_floats(a.batch, a.L.n_in, a.L.D) * sizeof(float);
    if (smem_ws <= 113 * 1024) return a.batch <= 24 ? launch_ws<3>(a, f, st) : launch_ws<4>(a, f, st);
  }
  switch ((a.batch + 7) / 8) {
    case 1: return launch_fused<1>(a, f, st);
    case 2: return launch_fused<2>(a, f, st);
    case 3: return launch_fused<3>(a, f, st);
    default: return launch_fused<4>(a, f, st);
  }
}

}  // namespace sgmcmc
