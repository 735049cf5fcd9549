// K5r: BNN-SGHMC with the chain RESIDENT on an SM for a whole block of steps.
//
// `sample, cost = next(sampler)` of the BOHAMIANN configuration is, per chain, the cost + gradient of
// pysgmcmc/models/bayesian_neural_network.py:28-69,337-388 followed by the update of
// pysgmcmc/samplers/sghmc.py:165-251 -- and a chain never looks at another chain.  K4 + K1
// (bnn.cu, update_kernels.cu) walk ALL chains once per step: the right shape for thousands of chains (every
// step streams the state through HBM at the roofline), the wrong one for the reference's own use of the
// sampler -- ONE chain (bayesian_neural_network.py:436-531), where a step is two launches of two-warp CTAs.
// Here one CTA owns one chain for `n_steps` steps: theta, V, tau, g, v_hat and minv (7 x D floats = 147 KB
// for D = 5252) are loaded into shared memory once, the minibatch rows of the next step arrive by cp.async
// while the current step computes, and only costs and thinned samples leave.
//
// Arithmetic.  The cost + gradient is the FFMA formulation of bnn.cu (variant 0) spread over more threads:
// every dot product is accumulated in the same order (bias first, k ascending; minibatch rows ascending for
// the weight gradients), so the gradient equals that kernel's bit for bit except d/d rho, d/d b4 (scalar
// expressions) and the cost's prior term (sum of theta^2 in another order).  The update is sampler_math.cuh's
// arithmetic with K1's (element group, step) -> Philox counter mapping: given the same gradient the new state
// is K1's (and the oracle's) bit for bit (tests/test_bnn_resident_gpu.py).
//
// Two kinds of warps, side by side (672 threads, one CTA per SM):
//   warps 0-7, the gradient: a step is eight phases separated by a named barrier of these 256 threads --
//     A  layer 1, aligned copy of W4        B, C  layers 2, 3          D  head: f, d cost / d f, squared errors
//     E  dZ3 | dW4                          F  dW3, db3 | dZ2          G  dW2, db2 | dZ1        H  dW1, db1
//     A GEMM's workers own 2 units x 4 minibatch rows (8 accumulators, one 64-bit weight load and one 128-bit
//     load of a TRANSPOSED activation copy [unit][row] per k) and are dealt to warps 0-3, one per SM
//     sub-partition; the second GEMM of a backward phase (no data dependence) runs on warps 4-7.
//   warps 8-20, the update: everything of the SGHMC step that does not need the gradient -- the Philox normals,
//     r_t, tau_t, minv_t = 1 / sqrt(v_hat), sigma * z (5 of the 8 IEEE divisions / square roots per element) --
//     WHILE the gradient warps compute; warp 20 then evaluates the scalar tail of the cost.
//   After one CTA barrier all 21 warps finish the update (g_t, v_hat_t, v_t, theta_t: multiplies and adds) and
//   sum theta^2 for the next step's prior.  When the two extra D-float arrays do not fit (minibatch > 20 rows)
//   the whole update runs after the barrier instead (sgmcmc_set_bnn_resident_overlap(0) forces that; same bits).
#include "bnn_common.cuh"

namespace sgmcmc {

constexpr int RS_T = 672;        // 21 warps: D/4 = 1313 update groups are 2 rounds of 672 (97.7 % of the slots)
constexpr int RS_GW = 8;         // warps 0 .. 7: the gradient (256 threads); warps 8 .. 20: the update's gradient-free part
constexpr int RS_GT = 32 * RS_GW;
constexpr int RS_UT = RS_T - RS_GT;
constexpr int RS_KG = 12;        // k values per worker of a weight-gradient GEMM (5 warps cover 50 + 2 padding)
constexpr int RS_NKG = 5;
constexpr int RS_HALF = HID / 2; // a gradient worker owns 2 units: (2 jj, 2 jj + 1) or (kk, kk + 25)
constexpr int RS_WT = 128;       // workers of a GEMM are dealt to 4 warps, one per SM sub-partition
constexpr int RS_MAX_BATCH = 32;

struct ResidentArgs {
  float *theta, *v, *tau, *g, *v_hat, *minv;      // [C, D]
  const float *X, *y;
  const int32_t* starts;                          // [n_steps, C] or NULL
  const float* z;                                 // [n_steps, C, D] or NULL
  float *trace, *cost_trace;                      // [n_steps / keep_every, C, D], [.., C] or NULL
  float *cost_last, *cost_all;                    // [C]; [n_steps, C] or NULL
  float* grad_out;                                // [C, D] or NULL: the gradient of the LAST step (tests)
  int64_t n_chains, n_steps, n_burn_in, keep_every;
  int batch, adapt_forever, pre;
  float inv_bs, inv_n, prior_den_inv;
  BnnLayout L;
  SghmcScalars<float> s;
  uint64_t seed, step0, group_offset;
};

struct ResidentSmem {
  float *TH, *G, *V, *TAU, *GG, *VH, *MINV;
  float *PS, *PR;                                 // pre mode: sigma * z and r_t of the step (else NULL)
  float *H1, *H2, *H3, *Z3, *Z2, *Z1;             // [BP][HS]
  float *H1t, *H2t, *Z3t, *Z2t;                   // [HID][TS]
  float *X0, *Y0;                                 // two buffers each, XB / YB floats apart
  int XB, YB;
  float *sDf, *sSe, *sW4, *scr;
  int BP, TS, total, zero_from;
};

// rows padded to a multiple of 4; the transposed buffers' row stride TS has TS / 4 odd, so that the 128-bit
// accesses of 8 consecutive units fall into 8 different 16-byte bank groups
__host__ __device__ inline ResidentSmem resident_carve(float* base, int batch, int n_in, int D, int pre) {
  ResidentSmem s;
  s.BP = (batch + 3) & ~3;
  s.TS = ((s.BP >> 2) & 1) ? s.BP : s.BP + 4;
  const int RB = s.BP * HS, TB = HID * s.TS, XB = (s.BP * n_in + 3) & ~3, YB = s.BP;
  float* p = base;
  s.TH = p; p += D; s.G = p; p += D; s.V = p; p += D; s.TAU = p; p += D;
  s.GG = p; p += D; s.VH = p; p += D; s.MINV = p; p += D;
  s.PS = s.PR = nullptr;
  if (pre) { s.PS = p; p += D; s.PR = p; p += D; }
  s.zero_from = (int)(p - base);
  s.H1 = p; p += RB; s.H2 = p; p += RB; s.H3 = p; p += RB;
  s.Z3 = p; p += RB; s.Z2 = p; p += RB; s.Z1 = p; p += RB;
  s.H1t = p; p += TB; s.H2t = p; p += TB; s.Z3t = p; p += TB; s.Z2t = p; p += TB;
  s.X0 = p; p += 2 * XB;
  s.Y0 = p; p += 2 * YB;
  s.XB = XB; s.YB = YB;
  s.sDf = p; p += RS_MAX_BATCH; s.sSe = p; p += RS_MAX_BATCH; s.sW4 = p; p += 64; s.scr = p; p += 64;
  s.total = (int)(p - base);
  return s;
}

__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)),
               "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// named barriers: 1 = the gradient warps among themselves, 2 = "head done" from the gradient warps to warp 20
__device__ __forceinline__ void grad_sync() { asm volatile("bar.sync 1, %0;" ::"n"(RS_GT) : "memory"); }
__device__ __forceinline__ void head_done_arrive() {
  __threadfence_block();
  asm volatile("bar.arrive 2, %0;" ::"n"(RS_GT + 32) : "memory");
}
__device__ __forceinline__ void head_done_wait() { asm volatile("bar.sync 2, %0;" ::"n"(RS_GT + 32) : "memory"); }

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }

// rows i0 .. i0+3 of a tanh layer for units 2 jj, 2 jj + 1: out = tanh(b + sum_k in[.][k] W[k][.]), k ascending
__device__ __forceinline__ void rs_forward(const float* __restrict__ TH, int oW, int ob,
                                           const float* __restrict__ inT, float* __restrict__ outR,
                                           float* __restrict__ outT, const int TS, int i0, int jj) {
  const float2 b = ld2(TH + ob + 2 * jj);
  float a0[4], a1[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) { a0[r] = b.x; a1[r] = b.y; }
  const float* w = TH + oW + 2 * jj;
  const float* h = inT + i0;
#pragma unroll 10
  for (int k = 0; k < HID; ++k) {
    const float2 wv = ld2(w + k * HID);
    const float4 hv = ld4(h + k * TS);
    a0[0] = fmaf(hv.x, wv.x, a0[0]); a0[1] = fmaf(hv.y, wv.x, a0[1]); a0[2] = fmaf(hv.z, wv.x, a0[2]); a0[3] = fmaf(hv.w, wv.x, a0[3]);
    a1[0] = fmaf(hv.x, wv.y, a1[0]); a1[1] = fmaf(hv.y, wv.y, a1[1]); a1[2] = fmaf(hv.z, wv.y, a1[2]); a1[3] = fmaf(hv.w, wv.y, a1[3]);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    a0[r] = fast_tanh(a0[r]); a1[r] = fast_tanh(a1[r]);
    st2(outR + (i0 + r) * HS + 2 * jj, a0[r], a1[r]);
  }
  if (outT != nullptr) {
    st4(outT + (2 * jj) * TS + i0, a0[0], a0[1], a0[2], a0[3]);
    st4(outT + (2 * jj + 1) * TS + i0, a1[0], a1[1], a1[2], a1[3]);
  }
}

// rows i0 .. i0+3 of dZ_{l-1}[.][k] = (sum_m dZ_l[.][m] W_l[k][m]) * (1 - H_{l-1}[.][k]^2) for k = kk and kk + 25,
// m ascending
__device__ __forceinline__ void rs_backward_data(const float* __restrict__ TH, int oW,
                                                 const float* __restrict__ zT, const float* __restrict__ hT,
                                                 float* __restrict__ outR, float* __restrict__ outT, const int TS,
                                                 int i0, int kk) {
  const int k0 = kk, k1 = kk + RS_HALF;
  const float* w0 = TH + oW + k0 * HID;
  const float* w1 = TH + oW + k1 * HID;
  const float* zt = zT + i0;
  float a0[4] = {0.0f, 0.0f, 0.0f, 0.0f}, a1[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 5
  for (int m2 = 0; m2 < HID / 2; ++m2) {
    const float2 wa = ld2(w0 + 2 * m2), wb = ld2(w1 + 2 * m2);
    const float4 z0 = ld4(zt + (2 * m2) * TS), z1 = ld4(zt + (2 * m2 + 1) * TS);
    a0[0] = fmaf(z0.x, wa.x, a0[0]); a0[1] = fmaf(z0.y, wa.x, a0[1]); a0[2] = fmaf(z0.z, wa.x, a0[2]); a0[3] = fmaf(z0.w, wa.x, a0[3]);
    a1[0] = fmaf(z0.x, wb.x, a1[0]); a1[1] = fmaf(z0.y, wb.x, a1[1]); a1[2] = fmaf(z0.z, wb.x, a1[2]); a1[3] = fmaf(z0.w, wb.x, a1[3]);
    a0[0] = fmaf(z1.x, wa.y, a0[0]); a0[1] = fmaf(z1.y, wa.y, a0[1]); a0[2] = fmaf(z1.z, wa.y, a0[2]); a0[3] = fmaf(z1.w, wa.y, a0[3]);
    a1[0] = fmaf(z1.x, wb.y, a1[0]); a1[1] = fmaf(z1.y, wb.y, a1[1]); a1[2] = fmaf(z1.z, wb.y, a1[2]); a1[3] = fmaf(z1.w, wb.y, a1[3]);
  }
  const float4 h0 = ld4(hT + k0 * TS + i0), h1 = ld4(hT + k1 * TS + i0);
  a0[0] = a0[0] * fmaf(-h0.x, h0.x, 1.0f); a0[1] = a0[1] * fmaf(-h0.y, h0.y, 1.0f);
  a0[2] = a0[2] * fmaf(-h0.z, h0.z, 1.0f); a0[3] = a0[3] * fmaf(-h0.w, h0.w, 1.0f);
  a1[0] = a1[0] * fmaf(-h1.x, h1.x, 1.0f); a1[1] = a1[1] * fmaf(-h1.y, h1.y, 1.0f);
  a1[2] = a1[2] * fmaf(-h1.z, h1.z, 1.0f); a1[3] = a1[3] * fmaf(-h1.w, h1.w, 1.0f);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    outR[(i0 + r) * HS + k0] = a0[r];
    outR[(i0 + r) * HS + k1] = a1[r];
  }
  if (outT != nullptr) {
    st4(outT + k0 * TS + i0, a0[0], a0[1], a0[2], a0[3]);
    st4(outT + k1 * TS + i0, a1[0], a1[1], a1[2], a1[3]);
  }
}

// G[W_l[k][j]] = theta * pscale + sum_i H_{l-1}[i][k] dZ_l[i][j] for j = 2 jj, 2 jj + 1 and k = 12 kg .. 12 kg + 11
// (< 50), i ascending; kg == 0 also sums db[j]
template <int BATCH_CT>
__device__ __forceinline__ void rs_weight_grad(const float* __restrict__ TH, float* __restrict__ G, int oW, int ob,
                                               const float* __restrict__ hR, const float* __restrict__ zR,
                                               int batch_rt, float pscale, int kg, int jj) {
  const int batch = BATCH_CT > 0 ? BATCH_CT : batch_rt;
  float acc0[RS_KG], acc1[RS_KG], db0 = 0.0f, db1 = 0.0f;
#pragma unroll
  for (int kk = 0; kk < RS_KG; ++kk) { acc0[kk] = 0.0f; acc1[kk] = 0.0f; }
  const bool full = kg < RS_NKG - 1;               // the last group holds k = 48, 49 (and the padding 50, 51)
  const float* hp = hR + RS_KG * kg;
  const float* zp = zR + 2 * jj;
#pragma unroll 4
  for (int i = 0; i < batch; ++i) {
    const float2 d = ld2(zp + i * HS);
    db0 += d.x; db1 += d.y;
#pragma unroll
    for (int q = 0; q < 3; ++q)
      if (q == 0 || full) {
        const float4 h = ld4(hp + i * HS + 4 * q);
        acc0[4 * q + 0] = fmaf(h.x, d.x, acc0[4 * q + 0]); acc0[4 * q + 1] = fmaf(h.y, d.x, acc0[4 * q + 1]);
        acc0[4 * q + 2] = fmaf(h.z, d.x, acc0[4 * q + 2]); acc0[4 * q + 3] = fmaf(h.w, d.x, acc0[4 * q + 3]);
        acc1[4 * q + 0] = fmaf(h.x, d.y, acc1[4 * q + 0]); acc1[4 * q + 1] = fmaf(h.y, d.y, acc1[4 * q + 1]);
        acc1[4 * q + 2] = fmaf(h.z, d.y, acc1[4 * q + 2]); acc1[4 * q + 3] = fmaf(h.w, d.y, acc1[4 * q + 3]);
      }
  }
#pragma unroll
  for (int kk = 0; kk < RS_KG; ++kk) {
    const int k = RS_KG * kg + kk;
    if (k < HID) {
      const float2 t = ld2(TH + oW + k * HID + 2 * jj);
      st2(G + oW + k * HID + 2 * jj, fmaf(t.x, pscale, acc0[kk]), fmaf(t.y, pscale, acc1[kk]));
    }
  }
  if (kg == 0) {
    const float2 t = ld2(TH + ob + 2 * jj);
    st2(G + ob + 2 * jj, fmaf(t.x, pscale, db0), fmaf(t.y, pscale, db1));
  }
}

// Element group q of a chain row of D floats at `row` (D even: 8-byte aligned; D % 4 == 2: the last group holds two
// elements and odd chains start on an 8-byte boundary only, so those rows move as 64-bit halves).
__device__ __forceinline__ float4 rs_load_group(const float* __restrict__ row, int q, int D, float pad) {
  const float* p = row + 4 * q;
  if ((D & 3) == 0) return __ldcs(reinterpret_cast<const float4*>(p));
  const float2 lo = __ldcs(reinterpret_cast<const float2*>(p));
  if (4 * q + 2 < D) {
    const float2 hi = __ldcs(reinterpret_cast<const float2*>(p + 2));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
  }
  return make_float4(lo.x, lo.y, pad, pad);
}
__device__ __forceinline__ void rs_store_group(float* __restrict__ row, int q, int D, const float4& v) {
  float* p = row + 4 * q;
  if ((D & 3) == 0) { __stcs(reinterpret_cast<float4*>(p), v); return; }
  __stcs(reinterpret_cast<float2*>(p), make_float2(v.x, v.y));
  if (4 * q + 2 < D) __stcs(reinterpret_cast<float2*>(p + 2), make_float2(v.z, v.w));
}
// elements >= nv of a group are padding: keep their state at its neutral value
__device__ __forceinline__ void rs_fix_pad(float (&x)[4], int nv, float value) {
#pragma unroll
  for (int i = 2; i < 4; ++i)
    if (i >= nv) x[i] = value;
}

__device__ __forceinline__ void rs_unpack(const float4& q, float (&r)[4]) { r[0] = q.x; r[1] = q.y; r[2] = q.z; r[3] = q.w; }
__device__ __forceinline__ float4 rs_pack(const float (&r)[4]) { return make_float4(r[0], r[1], r[2], r[3]); }

// The update of sampler_math.cuh cut in two at the gradient: `pre` needs only the state (it runs on the update
// warps WHILE the gradient warps compute), `post` is what is left once the gradient exists.  Same operations,
// same order, same rounding as adapt() + sghmc_apply(): the tests compare with the oracle bit for bit.
__device__ __forceinline__ float rs_sample(float minv_t, float z, const SghmcScalars<float>& s) {
  using F = ieee<float>;
  const float noise_scale = F::sub(F::mul(s.a, minv_t), s.c4);                     // sghmc.py:211-217
  return F::mul(F::sqrt(F::max(noise_scale, 1e-16f)), z);                          // :220, base_classes.py:218
}
__device__ __forceinline__ void rs_adapt_pre(float& tau, float g, float v_hat, float& r_t, float& minv_t) {
  using F = ieee<float>;
  r_t = F::div(1.0f, F::add(tau, 1.0f));                                           // sghmc.py:168
  minv_t = safe_divide(1.0f, safe_sqrt(v_hat));                                    // :179-183
  tau = F::add(tau, F::add(safe_divide(F::mul(F::mul(-g, g), tau), v_hat), 1.0f)); // :172-176
}
__device__ __forceinline__ void rs_adapt_post(float& g, float& v_hat, float r_t, float grad) {
  using F = ieee<float>;
  const float g_t = F::add(g, F::add(F::mul(-r_t, g), F::mul(r_t, grad)));                          // :186-190
  v_hat = F::add(v_hat, F::add(F::mul(-r_t, v_hat), F::mul(r_t, F::mul(grad, grad))));              // :192-196
  g = g_t;
}
__device__ __forceinline__ void rs_apply_post(float& theta, float& v, float minv_t, float grad, float sample,
                                              const SghmcScalars<float>& s) {
  using F = ieee<float>;
  const float drift = F::sub(F::mul(F::mul(s.neg_eps2, minv_t), grad), F::mul(s.mdecay, v));
  const float v_t = F::add(v, F::add(drift, sample));                              // :233-238
  v = v_t;
  theta = F::add(theta, v_t);                                                      // :241-243
}

template <int BATCH_CT>
__global__ void __launch_bounds__(RS_T, 1) bnn_sghmc_resident_kernel(ResidentArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int NT = RS_T, NW = RS_T / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t chain = blockIdx.x;
  const BnnLayout L = a.L;
  const int D = L.D, n4 = (D + 3) >> 2, n_in = L.n_in;    // groups of 4 elements; D % 4 == 2: the last one is half padding
  const int batch = BATCH_CT > 0 ? BATCH_CT : a.batch;
  const bool pre = a.pre != 0;
  const ResidentSmem s = resident_carve(smem, batch, n_in, 4 * n4, a.pre);
  const int TS = s.TS, nrg = s.BP >> 2;
  const float pscale = a.prior_den_inv * a.inv_n;
  const int64_t g0 = chain * n4;                    // first element group of this chain (Philox counter, with group_offset)
  const int64_t row0 = chain * D;                   // the chain's row in the [C, D] arrays

  // ---- the chain's state -> shared memory; sum(theta^2); activation buffers and minibatch rows zeroed ----
  {
    float sq = 0.0f;
    for (int q = tid; q < n4; q += NT) {
      const float4 t = rs_load_group(a.theta + row0, q, D, 0.0f);
      sq = fmaf(t.x, t.x, sq); sq = fmaf(t.y, t.y, sq); sq = fmaf(t.z, t.z, sq); sq = fmaf(t.w, t.w, sq);
      reinterpret_cast<float4*>(s.TH)[q] = t;
      reinterpret_cast<float4*>(s.V)[q] = rs_load_group(a.v + row0, q, D, 0.0f);
      reinterpret_cast<float4*>(s.TAU)[q] = rs_load_group(a.tau + row0, q, D, 1.0f);
      reinterpret_cast<float4*>(s.GG)[q] = rs_load_group(a.g + row0, q, D, 1.0f);
      reinterpret_cast<float4*>(s.VH)[q] = rs_load_group(a.v_hat + row0, q, D, 1.0f);
      reinterpret_cast<float4*>(s.MINV)[q] = rs_load_group(a.minv + row0, q, D, 1.0f);
      reinterpret_cast<float4*>(s.G)[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);     // (the padding's gradient stays 0)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (lane == 0) s.scr[warp] = sq;
    for (int q = s.zero_from + tid; q < s.total - 64; q += NT) smem[q] = 0.0f;      // everything but scr
  }
  __syncthreads();
  auto start_of = [&](int64_t step) -> int64_t {
    return a.starts != nullptr ? (int64_t)__ldg(a.starts + step * a.n_chains + chain) : 0;
  };
  auto fetch_rows = [&](int64_t start, int buf) {    // X[start : start + B], y[start : start + B] -> buffer `buf`
    const float* xs = a.X + start * n_in;
    const float* ys = a.y + start;
    for (int t = tid; t < batch * n_in; t += RS_GT) cp_async4(s.X0 + buf * s.XB + t, xs + t);
    for (int t = tid; t < batch; t += RS_GT) cp_async4(s.Y0 + buf * s.YB + t, ys + t);
  };
  if (tid < RS_GT && a.n_steps > 0) fetch_rows(start_of(0), 0);

  for (int64_t st = 0; st < a.n_steps; ++st) {
    const int cur = (int)(st & 1);
    const float* sX = s.X0 + cur * s.XB;
    const float* sY = s.Y0 + cur * s.YB;
    const bool burn_in = a.adapt_forever || st < a.n_burn_in;
    const float* zrow = a.z != nullptr ? a.z + (st * a.n_chains + chain) * D : nullptr;
    cp_async_wait_all();
    __syncthreads();                                 // rows of this step landed; the last update is visible

    if (tid < RS_GT) {
      // =================== gradient warps: cost + gradient of this step into G ===================
      // a GEMM's workers (pair of units, group of 4 rows) are dealt to warps 0-3; the second GEMM of a backward
      // phase to warps 4-7: one (two) busy warps per SM sub-partition
      const int wid = tid & (RS_WT - 1);
      const bool first = tid < RS_WT;
      const int n_fw = RS_HALF * nrg;                // workers of a forward / backward-data GEMM
      // ---- A: start index of the next step (used in H); aligned copy of W4; layer 1 ----
      const int64_t start_next = st + 1 < a.n_steps ? start_of(st + 1) : 0;
      float fvi = 0.0f;
      if (tid < batch) fvi = __fdiv_rn(1.0f, expf(s.TH[L.orho]) + 1e-16f);      // :368 (used in D)
      if (warp == 4)
        for (int t = lane; t < 64; t += 32) s.sW4[t] = t < HID ? s.TH[L.oW4 + t] : 0.0f;
      if (first)
        for (int wk = wid; wk < n_fw; wk += RS_WT) {
          const int rg = wk / RS_HALF, jj = wk - rg * RS_HALF, i0 = 4 * rg;
          const float2 b = ld2(s.TH + L.ob1 + 2 * jj);
          float z0[4], z1[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) { z0[r] = b.x; z1[r] = b.y; }
          for (int m = 0; m < n_in; ++m) {
            const float2 w = ld2(s.TH + L.oW1 + m * HID + 2 * jj);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const float x = sX[(i0 + r) * n_in + m];
              z0[r] = fmaf(x, w.x, z0[r]); z1[r] = fmaf(x, w.y, z1[r]);
            }
          }
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            z0[r] = fast_tanh(z0[r]); z1[r] = fast_tanh(z1[r]);
            st2(s.H1 + (i0 + r) * HS + 2 * jj, z0[r], z1[r]);
          }
          st4(s.H1t + (2 * jj) * TS + i0, z0[0], z0[1], z0[2], z0[3]);
          st4(s.H1t + (2 * jj + 1) * TS + i0, z1[0], z1[1], z1[2], z1[3]);
        }
      grad_sync();
      // ---- B, C: layers 2 and 3 ----
      if (first)
        for (int wk = wid; wk < n_fw; wk += RS_WT) {
          const int rg = wk / RS_HALF;
          rs_forward(s.TH, L.oW2, L.ob2, s.H1t, s.H2, s.H2t, TS, 4 * rg, wk - rg * RS_HALF);
        }
      grad_sync();
      if (first)
        for (int wk = wid; wk < n_fw; wk += RS_WT) {
          const int rg = wk / RS_HALF;
          rs_forward(s.TH, L.oW3, L.ob3, s.H2t, s.H3, nullptr, TS, 4 * rg, wk - rg * RS_HALF);
        }
      grad_sync();
      // ---- D: head f_i = b4 + H3[i, :] . W4, one thread per row (k ascending) ----
      if (tid < batch) {
        const float* hr = s.H3 + tid * HS;
        float f = s.TH[L.ob4];
#pragma unroll
        for (int k4 = 0; k4 < K4S; ++k4) {
          const float4 h = ld4(hr + 4 * k4), w = ld4(s.sW4 + 4 * k4);
          f = fmaf(h.x, w.x, f); f = fmaf(h.y, w.y, f); f = fmaf(h.z, w.z, f); f = fmaf(h.w, w.w, f);
        }
        const float diff = sY[tid] - f;
        s.sDf[tid] = -(diff * fvi) * a.inv_bs;                                  // d cost / d f_i
        s.sSe[tid] = diff * diff;                                               // :370
      }
      head_done_arrive();                            // warp 20 computes the scalar tail of the cost from here on
      grad_sync();
      // ---- E: dZ3 = (df W4^T) * (1 - H3^2) | dW4 ----
      if (first)
        for (int wk = wid; wk < n_fw; wk += RS_WT) {
          const int rg = wk / RS_HALF, jj = wk - rg * RS_HALF, i0 = 4 * rg;
          const float2 w4 = ld2(s.sW4 + 2 * jj);
          float o0[4], o1[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const float2 h = ld2(s.H3 + (i0 + r) * HS + 2 * jj);
            const float df = s.sDf[i0 + r];
            o0[r] = (df * w4.x) * fmaf(-h.x, h.x, 1.0f);
            o1[r] = (df * w4.y) * fmaf(-h.y, h.y, 1.0f);
            st2(s.Z3 + (i0 + r) * HS + 2 * jj, o0[r], o1[r]);
          }
          st4(s.Z3t + (2 * jj) * TS + i0, o0[0], o0[1], o0[2], o0[3]);
          st4(s.Z3t + (2 * jj + 1) * TS + i0, o1[0], o1[1], o1[2], o1[3]);
        }
      if (warp == 4 && lane < RS_HALF) {
        const int jj = lane;
        const float2 w4 = ld2(s.sW4 + 2 * jj);
        float dw0 = 0.0f, dw1 = 0.0f;
        for (int i = 0; i < batch; ++i) {
          const float2 h = ld2(s.H3 + i * HS + 2 * jj);
          const float df = s.sDf[i];
          dw0 = fmaf(h.x, df, dw0); dw1 = fmaf(h.y, df, dw1);
        }
        st2(s.G + L.oW4 + 2 * jj, fmaf(w4.x, pscale, dw0), fmaf(w4.y, pscale, dw1));
      }
      grad_sync();
      // ---- F: dW3, db3 (warps 0-3) | dZ2 (warps 4-7);   G: dW2, db2 | dZ1 ----
      if (first) {
        if (wid < RS_HALF * RS_NKG) {
          const int kg = wid / RS_HALF;
          rs_weight_grad<BATCH_CT>(s.TH, s.G, L.oW3, L.ob3, s.H2, s.Z3, batch, pscale, kg, wid - kg * RS_HALF);
        }
      } else {
        for (int wk = wid; wk < n_fw; wk += RS_WT) {
          const int rg = wk / RS_HALF;
          rs_backward_data(s.TH, L.oW3, s.Z3t, s.H2t, s.Z2, s.Z2t, TS, 4 * rg, wk - rg * RS_HALF);
        }
      }
      grad_sync();
      if (first) {
        if (wid < RS_HALF * RS_NKG) {
          const int kg = wid / RS_HALF;
          rs_weight_grad<BATCH_CT>(s.TH, s.G, L.oW2, L.ob2, s.H1, s.Z2, batch, pscale, kg, wid - kg * RS_HALF);
        }
      } else {
        for (int wk = wid; wk < n_fw; wk += RS_WT) {
          const int rg = wk / RS_HALF;
          rs_backward_data(s.TH, L.oW2, s.Z2t, s.H1t, s.Z1, nullptr, TS, 4 * rg, wk - rg * RS_HALF);
        }
      }
      grad_sync();
      // ---- H: db1[j] = sum_i dZ1[i][j];  dW1[m][j] = sum_i X[i][m] dZ1[i][j];  rows of the next step ----
      if (st + 1 < a.n_steps) fetch_rows(start_next, cur ^ 1);
      for (int item = tid; item < (n_in + 1) * HID; item += RS_GT) {
        const int m = item / HID, j = item - m * HID;
        if (m == n_in) {
          float db = 0.0f;
          for (int i = 0; i < batch; ++i) db += s.Z1[i * HS + j];
          s.G[L.ob1 + j] = fmaf(s.TH[L.ob1 + j], pscale, db);
        } else {
          float dw = 0.0f;
          for (int i = 0; i < batch; ++i) dw = fmaf(sX[i * n_in + m], s.Z1[i * HS + j], dw);
          s.G[L.oW1 + m * HID + j] = fmaf(s.TH[L.oW1 + m * HID + j], pscale, dw);
        }
      }
    } else {
      // =================== update warps: everything of the update that does not need the gradient ===================
      if (pre) {
        for (int q = tid - RS_GT; q < n4; q += RS_UT) {
          float zf[4], ps[4];
          const int nv = min(4, D - 4 * q);
          if (zrow != nullptr) rs_unpack(rs_load_group(zrow, q, D, 0.0f), zf);
          else normal4((uint64_t)(g0 + q) + a.group_offset, a.step0 + (uint64_t)st, a.seed, zf);
          if (burn_in) {
            float ta[4], g[4], h[4], r[4], mi[4];
            rs_unpack(reinterpret_cast<const float4*>(s.TAU)[q], ta);
            rs_unpack(reinterpret_cast<const float4*>(s.GG)[q], g);
            rs_unpack(reinterpret_cast<const float4*>(s.VH)[q], h);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              rs_adapt_pre(ta[i], g[i], h[i], r[i], mi[i]);
              ps[i] = rs_sample(mi[i], zf[i], a.s);
            }
            rs_fix_pad(ta, nv, 1.0f); rs_fix_pad(mi, nv, 1.0f);
            reinterpret_cast<float4*>(s.TAU)[q] = rs_pack(ta);
            reinterpret_cast<float4*>(s.PR)[q] = rs_pack(r);
            reinterpret_cast<float4*>(s.MINV)[q] = rs_pack(mi);     // the inverse mass matrix of this step
          } else {
            float mi[4];
            rs_unpack(reinterpret_cast<const float4*>(s.MINV)[q], mi);
#pragma unroll
            for (int i = 0; i < 4; ++i) ps[i] = rs_sample(mi[i], zf[i], a.s);
          }
          reinterpret_cast<float4*>(s.PS)[q] = rs_pack(ps);
        }
      }
      if (warp == NW - 1) {
        // ---- the scalar tail of the cost (:372-388), d/d rho, d/d b4: one thread, in the shadow of E .. H ----
        head_done_wait();
        if (lane == 0) {
          const float rho = s.TH[L.orho], b4 = s.TH[L.ob4];
          const float e_rho = expf(rho);
          const float fvi = __fdiv_rn(1.0f, e_rho + 1e-16f);
          float sse = 0.0f, sdf = 0.0f, sq_t = 0.0f;
          for (int i = 0; i < batch; ++i) { sse += s.sSe[i]; sdf += s.sDf[i]; }
          for (int w = 0; w < NW; ++w) sq_t += s.scr[w];
          const float lv_den = 0.02f + 3e-16f;                              // safe_divide(., 2 * var)
          const float dl = rho - logf(1e-6f);
          const float nb = (float)batch;
          const float log_like_data = __fmul_rn(__fsub_rn(__fmul_rn(-sse, __fmul_rn(0.5f, fvi)),
                                                          __fmul_rn(__fmul_rn(0.5f, rho), nb)), a.inv_bs);
          const float lv = __fsub_rn(__fdiv_rn(-__fmul_rn(dl, dl), lv_den), 0.5f * logf(0.01f));     // :102-107
          const float wp = __fmul_rn(__fmul_rn(-0.5f, sq_t), a.prior_den_inv);                        // :131-141
          const float cost = -__fadd_rn(log_like_data, __fmul_rn(__fadd_rn(lv, wp), a.inv_n));
          const float drho_data = __fmul_rn(-__fsub_rn(__fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(0.5f, sse), e_rho), fvi), fvi),
                                                       __fmul_rn(0.5f, nb)), a.inv_bs);
          s.G[L.orho] = __fadd_rn(__fadd_rn(drho_data, __fmul_rn(__fdiv_rn(__fmul_rn(2.0f, dl), lv_den), a.inv_n)),
                                  __fmul_rn(rho, pscale));
          s.G[L.ob4] = fmaf(b4, pscale, sdf);
          // the cost of this step (at the parameters BEFORE its update, base_classes.py:258-310)
          if (a.cost_all != nullptr) a.cost_all[st * a.n_chains + chain] = cost;
          if (st == a.n_steps - 1) a.cost_last[chain] = cost;
          if (a.cost_trace != nullptr && (st + 1) % a.keep_every == 0)
            a.cost_trace[((st + 1) / a.keep_every - 1) * a.n_chains + chain] = cost;
        }
      }
    }
    __syncthreads();                                 // the gradient is complete; so is the gradient-free part
    // ---- the rest of the SGHMC update, all warps (K1's element group -> Philox counter mapping) ----
    {
      const bool store_minv = burn_in && (st == a.n_burn_in - 1 || (a.adapt_forever && st == a.n_steps - 1));
      const bool snap = a.trace != nullptr && (st + 1) % a.keep_every == 0;
      float* trow = snap ? a.trace + (((st + 1) / a.keep_every - 1) * a.n_chains + chain) * D : nullptr;
      if (a.grad_out != nullptr && st == a.n_steps - 1)
        for (int q = tid; q < n4; q += NT) rs_store_group(a.grad_out + row0, q, D, reinterpret_cast<const float4*>(s.G)[q]);
      float sq = 0.0f;
      for (int q = tid; q < n4; q += NT) {
        float t[4], v[4], gr[4];
        const int nv = min(4, D - 4 * q);
        rs_unpack(reinterpret_cast<const float4*>(s.G)[q], gr);
        rs_unpack(reinterpret_cast<const float4*>(s.TH)[q], t);
        rs_unpack(reinterpret_cast<const float4*>(s.V)[q], v);
        if (pre) {
          float ps[4], mi[4];
          rs_unpack(reinterpret_cast<const float4*>(s.PS)[q], ps);
          rs_unpack(reinterpret_cast<const float4*>(s.MINV)[q], mi);
          if (burn_in) {
            float g[4], h[4], r[4];
            rs_unpack(reinterpret_cast<const float4*>(s.GG)[q], g);
            rs_unpack(reinterpret_cast<const float4*>(s.VH)[q], h);
            rs_unpack(reinterpret_cast<const float4*>(s.PR)[q], r);
#pragma unroll
            for (int i = 0; i < 4; ++i) rs_adapt_post(g[i], h[i], r[i], gr[i]);
            rs_fix_pad(g, nv, 1.0f); rs_fix_pad(h, nv, 1.0f);
            reinterpret_cast<float4*>(s.GG)[q] = rs_pack(g);
            reinterpret_cast<float4*>(s.VH)[q] = rs_pack(h);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) rs_apply_post(t[i], v[i], mi[i], gr[i], ps[i], a.s);
        } else {
          float zf[4], mi[4];
          if (zrow != nullptr) rs_unpack(rs_load_group(zrow, q, D, 0.0f), zf);
          else normal4((uint64_t)(g0 + q) + a.group_offset, a.step0 + (uint64_t)st, a.seed, zf);
          if (burn_in) {
            float ta[4], g[4], h[4];
            rs_unpack(reinterpret_cast<const float4*>(s.TAU)[q], ta);
            rs_unpack(reinterpret_cast<const float4*>(s.GG)[q], g);
            rs_unpack(reinterpret_cast<const float4*>(s.VH)[q], h);
#pragma unroll
            for (int i = 0; i < 4; ++i) mi[i] = adapt(ta[i], g[i], h[i], gr[i]);
            rs_fix_pad(ta, nv, 1.0f); rs_fix_pad(g, nv, 1.0f); rs_fix_pad(h, nv, 1.0f); rs_fix_pad(mi, nv, 1.0f);
            reinterpret_cast<float4*>(s.TAU)[q] = rs_pack(ta);
            reinterpret_cast<float4*>(s.GG)[q] = rs_pack(g);
            reinterpret_cast<float4*>(s.VH)[q] = rs_pack(h);
            if (store_minv) reinterpret_cast<float4*>(s.MINV)[q] = rs_pack(mi);
          } else {
            rs_unpack(reinterpret_cast<const float4*>(s.MINV)[q], mi);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) sghmc_apply(t[i], v[i], mi[i], gr[i], zf[i], a.s);
        }
        rs_fix_pad(t, nv, 0.0f); rs_fix_pad(v, nv, 0.0f);
        sq = fmaf(t[0], t[0], sq); sq = fmaf(t[1], t[1], sq); sq = fmaf(t[2], t[2], sq); sq = fmaf(t[3], t[3], sq);
        const float4 tn = rs_pack(t);
        reinterpret_cast<float4*>(s.TH)[q] = tn;
        reinterpret_cast<float4*>(s.V)[q] = rs_pack(v);
        if (snap) rs_store_group(trow, q, D, tn);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      if (lane == 0) s.scr[warp] = sq;               // sum(theta^2) of the next step's cost
    }
  }
  __syncthreads();
  // ---- the state goes back ----
  {
    for (int q = tid; q < n4; q += NT) {
      rs_store_group(a.theta + row0, q, D, reinterpret_cast<const float4*>(s.TH)[q]);
      rs_store_group(a.v + row0, q, D, reinterpret_cast<const float4*>(s.V)[q]);
      rs_store_group(a.tau + row0, q, D, reinterpret_cast<const float4*>(s.TAU)[q]);
      rs_store_group(a.g + row0, q, D, reinterpret_cast<const float4*>(s.GG)[q]);
      rs_store_group(a.v_hat + row0, q, D, reinterpret_cast<const float4*>(s.VH)[q]);
      rs_store_group(a.minv + row0, q, D, reinterpret_cast<const float4*>(s.MINV)[q]);
    }
  }
}

constexpr size_t RS_SMEM_MAX = 227 * 1024;

// 1: the gradient-free part of the update runs beside the gradient (two more D-float arrays); 0: it does not fit
static int resident_pre_mode(int n_in, int batch) {
  const int Dp = (make_layout(n_in).D + 3) & ~3;
  return (size_t)resident_carve(nullptr, batch, n_in, Dp, 1).total * sizeof(float) <= RS_SMEM_MAX ? 1 : 0;
}

template <int BATCH_CT>
static void launch_resident(const ResidentArgs& a, size_t smem, cudaStream_t st) {
  auto k = bnn_sghmc_resident_kernel<BATCH_CT>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<(unsigned)a.n_chains, RS_T, smem, st>>>(a);
}

static bool resident_shape_ok(int n_in, int batch) {
  if (n_in < 1 || n_in > 64 || batch < 1 || batch > RS_MAX_BATCH) return false;
  const int Dp = (make_layout(n_in).D + 3) & ~3;    // D is even; D % 4 == 2: every chain's last element group is half padding
  return (size_t)resident_carve(nullptr, batch, n_in, Dp, 0).total * sizeof(float) <= RS_SMEM_MAX;
}

}  // namespace sgmcmc

using namespace sgmcmc;

static int g_resident_overlap = 1;

extern "C" int sgmcmc_set_bnn_resident_overlap(int on) {
  g_resident_overlap = on != 0;
  return SGMCMC_OK;
}

extern "C" int sgmcmc_bnn_resident_supported(int n_in, int batch) { return resident_shape_ok(n_in, batch) ? 1 : 0; }

extern "C" int sgmcmc_bnn_sghmc_run_resident_f32(float* theta, float* v, float* tau, float* g, float* v_hat,
                                                 float* minv, const float* X, const float* y,
                                                 const int32_t* starts, const float* z, float* trace,
                                                 float* cost_trace, float* cost_all, float* cost_last,
                                                 float* grad_out, int64_t n_chains, int n_in, int batch,
                                                 float batch_size_cfg, int64_t n_examples, int64_t n_steps,
                                                 int64_t n_burn_in, int adapt_forever, int64_t keep_every,
                                                 float epsilon, float mdecay, float scale_grad, uint64_t seed,
                                                 uint64_t step0, uint64_t chain_offset, void* stream) {
  SG_REQUIRE(n_chains >= 0 && n_steps >= 0 && n_burn_in >= 0 && keep_every >= 1, SGMCMC_E_INVALID,
             "bnn_sghmc_run_resident: n_chains, n_steps, n_burn_in must be >= 0 and keep_every >= 1");
  SG_REQUIRE(theta && v && tau && g && v_hat && minv && X && y && cost_last, SGMCMC_E_INVALID,
             "bnn_sghmc_run_resident: state arrays, X, y and cost_last must not be NULL");
  SG_REQUIRE(resident_shape_ok(n_in, batch), SGMCMC_E_UNSUPPORTED,
             "bnn_sghmc_run_resident: n_in = %d (<= 64) / batch = %d (<= %d) does not fit an SM's shared memory; use "
             "sgmcmc_bnn_sghmc_run_f32", n_in, batch, RS_MAX_BATCH);
  SG_REQUIRE(batch_size_cfg > 0 && n_examples >= 1 && scale_grad > 0, SGMCMC_E_INVALID,
             "bnn_sghmc_run_resident: batch_size_cfg, n_examples and scale_grad must be > 0");
  ResidentArgs a;
  a.theta = theta; a.v = v; a.tau = tau; a.g = g; a.v_hat = v_hat; a.minv = minv;
  a.X = X; a.y = y; a.starts = starts; a.z = z; a.trace = trace; a.cost_trace = cost_trace;
  a.cost_last = cost_last; a.cost_all = cost_all; a.grad_out = grad_out;
  a.n_chains = n_chains; a.n_steps = n_steps; a.n_burn_in = n_burn_in; a.keep_every = keep_every;
  a.batch = batch; a.adapt_forever = adapt_forever;
  a.L = make_layout(n_in);
  a.inv_bs = 1.0f / batch_size_cfg;
  a.inv_n = 1.0f / (float)n_examples;
  a.prior_den_inv = 1.0f / ((float)a.L.D + 3e-16f);
  a.s = make_sghmc_scalars<float>(epsilon, mdecay, scale_grad);
  a.seed = seed; a.step0 = step0;
  // Philox counter of element group q of chain c: (chain_offset + c) * ceil(D / 4) + q -- K1's mapping when D % 4 == 0
  a.group_offset = chain_offset * (uint64_t)((a.L.D + 3) / 4);
  const size_t al = a.L.D % 4 == 0 ? 16 : 8;       // rows move as 128-bit groups, or as 64-bit halves when D % 4 == 2
  for (float* p : {theta, v, tau, g, v_hat, minv})
    SG_REQUIRE(aligned_to(p, al), SGMCMC_E_ALIGN, "bnn_sghmc_run_resident: state arrays must be %zu-byte aligned", al);
  SG_REQUIRE((!trace || aligned_to(trace, al)) && (!z || aligned_to(z, al)) && (!grad_out || aligned_to(grad_out, al)),
             SGMCMC_E_ALIGN, "bnn_sghmc_run_resident: trace, z and grad_out must be %zu-byte aligned", al);
  if (n_chains == 0 || n_steps == 0) return SGMCMC_OK;
  SG_REQUIRE(n_chains <= 0x7fffffff, SGMCMC_E_UNSUPPORTED, "bnn_sghmc_run_resident: at most 2^31 - 1 chains per launch");
  a.pre = g_resident_overlap ? resident_pre_mode(n_in, batch) : 0;
  const size_t smem = (size_t)resident_carve(nullptr, batch, n_in, (a.L.D + 3) & ~3, a.pre).total * sizeof(float);
  if (batch == 20) launch_resident<20>(a, smem, (cudaStream_t)stream);      // the reference's minibatch: strides fold
  else launch_resident<0>(a, smem, (cudaStream_t)stream);
  return check_launch("bnn_sghmc_resident_kernel");
}
