// K5r: BNN-SGHMC with the chain RESIDENT on an SM for a whole block of steps.
//
// `sample, cost = next(sampler)` of the BOHAMIANN configuration is, per chain, the cost + gradient of
// pysgmcmc/models/bayesian_neural_network.py:28-69,337-388 followed by the update of
// pysgmcmc/samplers/sghmc.py:165-251 -- and a chain never looks at another chain.  K4 + K1
// (bnn.cu, update_kernels.cu) walk ALL chains once per step, so every step streams the whole state
// through HBM (44 B per element during burn-in) and a single chain pays two kernel launches per step.
// Here one CTA owns one chain for `n_steps` steps: theta, V, tau, g, v_hat and minv (7 x D floats =
// 147 KB for D = 5252) are loaded into shared memory once, the minibatch rows of the next step arrive
// by cp.async while the current step computes, and only costs and thinned samples leave.  HBM traffic
// is 2 x 147 KB per chain and LAUNCH instead of 231 KB per chain and STEP; a single chain steps at the
// latency of one SM instead of two launches.
//
// Arithmetic.  The cost + gradient is the FFMA formulation of bnn.cu (variant 0) spread over more
// threads: every dot product is accumulated in the same order (bias first, k ascending; minibatch
// rows ascending for the weight gradients), so the gradient equals that kernel's bit for bit except
// d/d rho and the cost's prior term (sum of theta^2 in another order).  The update is sampler_math.cuh
// with K1's (element group, step) -> Philox counter mapping, so given the same gradient the new state
// is K1's bit for bit.
//
// Threads.  A step is a sequence of barrier-separated phases on one CTA of RS_T threads:
//   A  sum(theta^2), W4 -> aligned copy, layer 1             B, C  layers 2, 3 (50 x rows/4 workers)
//   D  head: f, d cost / d f, squared errors                 E  dZ3 | dW4 | the scalar tail (one thread)
//   F  dW3, db3 | dZ2       G  dW2, db2 | dZ1                H  dW1, db1
//   I  SGHMC update of the D/4 element groups by all threads (+ snapshot of the thinned sample)
// A GEMM worker is (column j, group of 4 minibatch rows) in a 64-thread slot (50 active): the 4 rows'
// activations come from a TRANSPOSED copy [unit][row] with one broadcast 128-bit load per k, the weight
// column straight from the staged theta (consecutive j: conflict free).  The two halves of a backward
// phase have no data dependence and run on different slots at the same time.
#include "bnn_common.cuh"

namespace sgmcmc {

constexpr int RS_T = 672;        // 21 warps: D/4 = 1313 update groups are 2 rounds of 672 (97.7 % of the slots)
constexpr int RS_SLOT = 64;
constexpr int RS_KG = 12;        // k values per worker of a weight-gradient GEMM (5 slots cover 50 + 2 padding)
constexpr int RS_NKG = 5;
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
  int batch, adapt_forever;
  float inv_bs, inv_n, prior_den_inv;
  BnnLayout L;
  SghmcScalars<float> s;
  uint64_t seed, step0, group_offset;
};

struct ResidentSmem {
  float *TH, *G, *V, *TAU, *GG, *VH, *MINV;
  float *H1, *H2, *H3, *Z3, *Z2, *Z1;             // [BP][HS]
  float *H1t, *H2t, *Z3t, *Z2t;                   // [HID][TS]
  float *X0, *Y0;                                 // two buffers each, XB / YB floats apart
  int XB, YB;
  float *sDf, *sSe, *sW4, *scr;
  int BP, TS, total;
};

// rows padded to a multiple of 4; the transposed buffers' row stride TS has TS / 4 odd, so that the 128-bit
// accesses of 8 consecutive units fall into 8 different 16-byte bank groups
__host__ __device__ inline ResidentSmem resident_carve(float* base, int batch, int n_in, int D) {
  ResidentSmem s;
  s.BP = (batch + 3) & ~3;
  s.TS = ((s.BP >> 2) & 1) ? s.BP : s.BP + 4;
  const int RB = s.BP * HS, TB = HID * s.TS, XB = (s.BP * n_in + 3) & ~3, YB = s.BP;
  float* p = base;
  s.TH = p; p += D; s.G = p; p += D; s.V = p; p += D; s.TAU = p; p += D;
  s.GG = p; p += D; s.VH = p; p += D; s.MINV = p; p += D;
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

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}

// rows i0 .. i0+3 of a tanh layer for unit j: out = tanh(b[j] + sum_k in[.][k] W[k][j]), k ascending
__device__ __forceinline__ void rs_forward(const float* __restrict__ TH, int oW, int ob,
                                           const float* __restrict__ inT, float* __restrict__ outR,
                                           float* __restrict__ outT, int TS, int i0, int j) {
  const float b = TH[ob + j];
  float a0 = b, a1 = b, a2 = b, a3 = b;
  const float* w = TH + oW + j;
  const float* h = inT + i0;
#pragma unroll 10
  for (int k = 0; k < HID; ++k) {
    const float wv = w[k * HID];
    const float4 hv = ld4(h + k * TS);
    a0 = fmaf(hv.x, wv, a0); a1 = fmaf(hv.y, wv, a1); a2 = fmaf(hv.z, wv, a2); a3 = fmaf(hv.w, wv, a3);
  }
  a0 = fast_tanh(a0); a1 = fast_tanh(a1); a2 = fast_tanh(a2); a3 = fast_tanh(a3);
  outR[(i0 + 0) * HS + j] = a0; outR[(i0 + 1) * HS + j] = a1;
  outR[(i0 + 2) * HS + j] = a2; outR[(i0 + 3) * HS + j] = a3;
  if (outT != nullptr) st4(outT + j * TS + i0, a0, a1, a2, a3);
}

// rows i0 .. i0+3 of dZ_{l-1}[.][k] = (sum_m dZ_l[.][m] W_l[k][m]) * (1 - H_{l-1}[.][k]^2), m ascending
__device__ __forceinline__ void rs_backward_data(const float* __restrict__ TH, int oW,
                                                 const float* __restrict__ zT, const float* __restrict__ hT,
                                                 float* __restrict__ outR, float* __restrict__ outT, int TS,
                                                 int i0, int k) {
  const float2* wrow = reinterpret_cast<const float2*>(TH + oW + k * HID);
  const float* zt = zT + i0;
  float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll 5
  for (int m2 = 0; m2 < HID / 2; ++m2) {
    const float2 w = wrow[m2];
    const float4 z0 = ld4(zt + (2 * m2) * TS), z1 = ld4(zt + (2 * m2 + 1) * TS);
    a0 = fmaf(z0.x, w.x, a0); a1 = fmaf(z0.y, w.x, a1); a2 = fmaf(z0.z, w.x, a2); a3 = fmaf(z0.w, w.x, a3);
    a0 = fmaf(z1.x, w.y, a0); a1 = fmaf(z1.y, w.y, a1); a2 = fmaf(z1.z, w.y, a2); a3 = fmaf(z1.w, w.y, a3);
  }
  const float4 hv = ld4(hT + k * TS + i0);
  a0 = a0 * fmaf(-hv.x, hv.x, 1.0f); a1 = a1 * fmaf(-hv.y, hv.y, 1.0f);
  a2 = a2 * fmaf(-hv.z, hv.z, 1.0f); a3 = a3 * fmaf(-hv.w, hv.w, 1.0f);
  outR[(i0 + 0) * HS + k] = a0; outR[(i0 + 1) * HS + k] = a1;
  outR[(i0 + 2) * HS + k] = a2; outR[(i0 + 3) * HS + k] = a3;
  if (outT != nullptr) st4(outT + k * TS + i0, a0, a1, a2, a3);
}

// G[W_l[k][j]] = theta * pscale + sum_i H_{l-1}[i][k] dZ_l[i][j] for k = 12 kg .. 12 kg + 11 (< 50), i ascending;
// kg == 0 also sums db[j]
__device__ __forceinline__ void rs_weight_grad(const float* __restrict__ TH, float* __restrict__ G, int oW, int ob,
                                               const float* __restrict__ hR, const float* __restrict__ zR,
                                               int batch, float pscale, int kg, int j) {
  float acc[RS_KG], db = 0.0f;
#pragma unroll
  for (int kk = 0; kk < RS_KG; ++kk) acc[kk] = 0.0f;
  const int nq = kg < RS_NKG - 1 ? 3 : 1;          // the last group holds k = 48, 49 (and the padding 50, 51)
  const float* hp = hR + RS_KG * kg;
  for (int i = 0; i < batch; ++i) {
    const float d = zR[i * HS + j];
    db += d;
#pragma unroll
    for (int q = 0; q < 3; ++q)
      if (q < nq) {
        const float4 h = ld4(hp + i * HS + 4 * q);
        acc[4 * q + 0] = fmaf(h.x, d, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(h.y, d, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(h.z, d, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(h.w, d, acc[4 * q + 3]);
      }
  }
#pragma unroll
  for (int kk = 0; kk < RS_KG; ++kk) {
    const int k = RS_KG * kg + kk;
    if (k < HID) G[oW + k * HID + j] = fmaf(TH[oW + k * HID + j], pscale, acc[kk]);
  }
  if (kg == 0) G[ob + j] = fmaf(TH[ob + j], pscale, db);
}

__device__ __forceinline__ void rs_unpack(const float4& q, float (&r)[4]) { r[0] = q.x; r[1] = q.y; r[2] = q.z; r[3] = q.w; }

template <int NT>
__global__ void __launch_bounds__(NT, 1) bnn_sghmc_resident_kernel(ResidentArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int NW = NT / 32, NSLOT = NT / RS_SLOT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot = tid / RS_SLOT, j = tid % RS_SLOT;
  const int64_t chain = blockIdx.x;
  const BnnLayout L = a.L;
  const int D = L.D, n4 = D >> 2, batch = a.batch, n_in = L.n_in;
  const ResidentSmem s = resident_carve(smem, batch, n_in, D);
  const int TS = s.TS, nrg = s.BP >> 2;
  const float pscale = a.prior_den_inv * a.inv_n;
  const bool split = RS_NKG + nrg <= NSLOT;         // the halves of a backward phase on different slots
  const int dw4_slot = nrg < NSLOT ? nrg : 0;

  // ---- the chain's state -> shared memory; activation buffers and minibatch rows zeroed -------------------
  {
    const int64_t g0 = chain * n4;
    const float4 *t4 = reinterpret_cast<const float4*>(a.theta) + g0, *v4 = reinterpret_cast<const float4*>(a.v) + g0,
                 *ta4 = reinterpret_cast<const float4*>(a.tau) + g0, *g4 = reinterpret_cast<const float4*>(a.g) + g0,
                 *h4 = reinterpret_cast<const float4*>(a.v_hat) + g0, *m4 = reinterpret_cast<const float4*>(a.minv) + g0;
    for (int q = tid; q < n4; q += NT) {
      reinterpret_cast<float4*>(s.TH)[q] = __ldcs(t4 + q);
      reinterpret_cast<float4*>(s.V)[q] = __ldcs(v4 + q);
      reinterpret_cast<float4*>(s.TAU)[q] = __ldcs(ta4 + q);
      reinterpret_cast<float4*>(s.GG)[q] = __ldcs(g4 + q);
      reinterpret_cast<float4*>(s.VH)[q] = __ldcs(h4 + q);
      reinterpret_cast<float4*>(s.MINV)[q] = __ldcs(m4 + q);
    }
    for (int q = tid; q < s.total - 7 * D; q += NT) s.H1[q] = 0.0f;     // everything after the state arrays
  }
  __syncthreads();
  auto fetch_rows = [&](int64_t step, int buf) {     // X[start : start + B], y[start : start + B] of `step`
    const int64_t start = a.starts != nullptr ? (int64_t)a.starts[step * a.n_chains + chain] : 0;
    const float* xs = a.X + start * n_in;
    const float* ys = a.y + start;
    for (int t = tid; t < batch * n_in; t += NT) cp_async4(s.X0 + buf * s.XB + t, xs + t);
    for (int t = tid; t < batch; t += NT) cp_async4(s.Y0 + buf * s.YB + t, ys + t);
  };
  if (a.n_steps > 0) fetch_rows(0, 0);

  for (int64_t st = 0; st < a.n_steps; ++st) {
    const int cur = (int)(st & 1);
    const float* sX = s.X0 + cur * s.XB;
    const float* sY = s.Y0 + cur * s.YB;
    cp_async_wait_all();
    __syncthreads();                                 // rows of this step landed; theta of the last update visible

    // ---- A: rows of the next step; sum(theta^2); W4; layer 1 ----
    if (st + 1 < a.n_steps) fetch_rows(st + 1, cur ^ 1);
    {
      float sq = 0.0f;
      for (int q = tid; q < n4; q += NT) {
        const float4 t = reinterpret_cast<const float4*>(s.TH)[q];
        sq = fmaf(t.x, t.x, sq); sq = fmaf(t.y, t.y, sq); sq = fmaf(t.z, t.z, sq); sq = fmaf(t.w, t.w, sq);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      if (lane == 0) s.scr[warp] = sq;
    }
    if (tid >= NT - 64) {
      const int t = tid - (NT - 64);
      s.sW4[t] = t < HID ? s.TH[L.oW4 + t] : 0.0f;
    }
    if (j < HID) for (int rg = slot; rg < nrg; rg += NSLOT) {
      const int i0 = 4 * rg;
      const float b = s.TH[L.ob1 + j];
      float z0 = b, z1 = b, z2 = b, z3 = b;
      for (int m = 0; m < n_in; ++m) {
        const float w = s.TH[L.oW1 + m * HID + j];
        z0 = fmaf(sX[(i0 + 0) * n_in + m], w, z0); z1 = fmaf(sX[(i0 + 1) * n_in + m], w, z1);
        z2 = fmaf(sX[(i0 + 2) * n_in + m], w, z2); z3 = fmaf(sX[(i0 + 3) * n_in + m], w, z3);
      }
      z0 = fast_tanh(z0); z1 = fast_tanh(z1); z2 = fast_tanh(z2); z3 = fast_tanh(z3);
      s.H1[(i0 + 0) * HS + j] = z0; s.H1[(i0 + 1) * HS + j] = z1;
      s.H1[(i0 + 2) * HS + j] = z2; s.H1[(i0 + 3) * HS + j] = z3;
      st4(s.H1t + j * TS + i0, z0, z1, z2, z3);
    }
    __syncthreads();
    // ---- B, C: layers 2 and 3 ----
    if (j < HID)
      for (int rg = slot; rg < nrg; rg += NSLOT) rs_forward(s.TH, L.oW2, L.ob2, s.H1t, s.H2, s.H2t, TS, 4 * rg, j);
    __syncthreads();
    if (j < HID)
      for (int rg = slot; rg < nrg; rg += NSLOT) rs_forward(s.TH, L.oW3, L.ob3, s.H2t, s.H3, nullptr, TS, 4 * rg, j);
    __syncthreads();
    // ---- D: head f_i = b4 + H3[i, :] . W4, one thread per row (k ascending) ----
    const float rho = s.TH[L.orho], b4 = s.TH[L.ob4];
    const float e_rho = expf(rho);
    const float fvi = __fdiv_rn(1.0f, e_rho + 1e-16f);                  // :368
    if (tid < batch) {
      const float* hr = s.H3 + tid * HS;
      float f = b4;
#pragma unroll
      for (int k4 = 0; k4 < K4S; ++k4) {
        const float4 h = ld4(hr + 4 * k4), w = ld4(s.sW4 + 4 * k4);
        f = fmaf(h.x, w.x, f); f = fmaf(h.y, w.y, f); f = fmaf(h.z, w.z, f); f = fmaf(h.w, w.w, f);
      }
      const float diff = sY[tid] - f;
      s.sDf[tid] = -(diff * fvi) * a.inv_bs;                            // d cost / d f_i
      s.sSe[tid] = diff * diff;                                         // :370
    }
    __syncthreads();
    // ---- E: dZ3 = (df W4^T) * (1 - H3^2) | dW4 | scalar tail of the cost (:372-388), d/d rho, d/d b4 ----
    if (j < HID) for (int rg = slot; rg < nrg; rg += NSLOT) {
      const int i0 = 4 * rg;
      const float w4 = s.sW4[j];
      float o[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float h = s.H3[(i0 + r) * HS + j];
        o[r] = (s.sDf[i0 + r] * w4) * fmaf(-h, h, 1.0f);
        s.Z3[(i0 + r) * HS + j] = o[r];
      }
      st4(s.Z3t + j * TS + i0, o[0], o[1], o[2], o[3]);
    }
    if (slot == dw4_slot && j < HID) {
      float dw = 0.0f;
      for (int i = 0; i < batch; ++i) dw = fmaf(s.H3[i * HS + j], s.sDf[i], dw);
      s.G[L.oW4 + j] = fmaf(s.sW4[j], pscale, dw);
    }
    if (tid == NT - 1) {
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
    __syncthreads();
    // ---- F: dW3, db3 | dZ2;   G: dW2, db2 | dZ1 ----
    if (split) {
      if (slot < RS_NKG) { if (j < HID) rs_weight_grad(s.TH, s.G, L.oW3, L.ob3, s.H2, s.Z3, batch, pscale, slot, j); }
      else if (slot < RS_NKG + nrg && j < HID)
        rs_backward_data(s.TH, L.oW3, s.Z3t, s.H2t, s.Z2, s.Z2t, TS, 4 * (slot - RS_NKG), j);
    } else {
      if (slot < RS_NKG && j < HID) rs_weight_grad(s.TH, s.G, L.oW3, L.ob3, s.H2, s.Z3, batch, pscale, slot, j);
      if (j < HID)
        for (int rg = slot; rg < nrg; rg += NSLOT) rs_backward_data(s.TH, L.oW3, s.Z3t, s.H2t, s.Z2, s.Z2t, TS, 4 * rg, j);
    }
    __syncthreads();
    if (split) {
      if (slot < RS_NKG) { if (j < HID) rs_weight_grad(s.TH, s.G, L.oW2, L.ob2, s.H1, s.Z2, batch, pscale, slot, j); }
      else if (slot < RS_NKG + nrg && j < HID)
        rs_backward_data(s.TH, L.oW2, s.Z2t, s.H1t, s.Z1, nullptr, TS, 4 * (slot - RS_NKG), j);
    } else {
      if (slot < RS_NKG && j < HID) rs_weight_grad(s.TH, s.G, L.oW2, L.ob2, s.H1, s.Z2, batch, pscale, slot, j);
      if (j < HID)
        for (int rg = slot; rg < nrg; rg += NSLOT) rs_backward_data(s.TH, L.oW2, s.Z2t, s.H1t, s.Z1, nullptr, TS, 4 * rg, j);
    }
    __syncthreads();
    // ---- H: db1[j] = sum_i dZ1[i][j];  dW1[m][j] = sum_i X[i][m] dZ1[i][j] ----
    for (int item = tid; item < (n_in + 1) * HID; item += NT) {
      const int m = item / HID, jj = item - m * HID;
      if (m == n_in) {
        float db = 0.0f;
        for (int i = 0; i < batch; ++i) db += s.Z1[i * HS + jj];
        s.G[L.ob1 + jj] = fmaf(s.TH[L.ob1 + jj], pscale, db);
      } else {
        float dw = 0.0f;
        for (int i = 0; i < batch; ++i) dw = fmaf(sX[i * n_in + m], s.Z1[i * HS + jj], dw);
        s.G[L.oW1 + m * HID + jj] = fmaf(s.TH[L.oW1 + m * HID + jj], pscale, dw);
      }
    }
    __syncthreads();
    // ---- I: the SGHMC update (sampler_math.cuh; K1's element group -> Philox counter mapping) ----
    {
      const bool burn_in = a.adapt_forever || st < a.n_burn_in;
      const bool store_minv = burn_in && (st == a.n_burn_in - 1 || (a.adapt_forever && st == a.n_steps - 1));
      const bool snap = a.trace != nullptr && (st + 1) % a.keep_every == 0;
      const int64_t g0 = chain * n4;
      const float4* z4 = a.z != nullptr ? reinterpret_cast<const float4*>(a.z) + (st * a.n_chains + chain) * n4 : nullptr;
      float4* tr4 = snap ? reinterpret_cast<float4*>(a.trace) + (((st + 1) / a.keep_every - 1) * a.n_chains + chain) * n4
                         : nullptr;
      if (a.grad_out != nullptr && st == a.n_steps - 1)
        for (int q = tid; q < n4; q += NT)
          reinterpret_cast<float4*>(a.grad_out)[g0 + q] = reinterpret_cast<const float4*>(s.G)[q];
      for (int q = tid; q < n4; q += NT) {
        float zf[4], t[4], v[4], gr[4];
        if (z4 != nullptr) rs_unpack(__ldcs(z4 + q), zf);
        else normal4((uint64_t)(g0 + q) + a.group_offset, a.step0 + (uint64_t)st, a.seed, zf);
        rs_unpack(reinterpret_cast<const float4*>(s.G)[q], gr);
        rs_unpack(reinterpret_cast<const float4*>(s.TH)[q], t);
        rs_unpack(reinterpret_cast<const float4*>(s.V)[q], v);
        if (burn_in) {
          float ta[4], g[4], h[4], mi[4];
          rs_unpack(reinterpret_cast<const float4*>(s.TAU)[q], ta);
          rs_unpack(reinterpret_cast<const float4*>(s.GG)[q], g);
          rs_unpack(reinterpret_cast<const float4*>(s.VH)[q], h);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            mi[i] = adapt(ta[i], g[i], h[i], gr[i]);
            sghmc_apply(t[i], v[i], mi[i], gr[i], zf[i], a.s);
          }
          reinterpret_cast<float4*>(s.TAU)[q] = make_float4(ta[0], ta[1], ta[2], ta[3]);
          reinterpret_cast<float4*>(s.GG)[q] = make_float4(g[0], g[1], g[2], g[3]);
          reinterpret_cast<float4*>(s.VH)[q] = make_float4(h[0], h[1], h[2], h[3]);
          if (store_minv) reinterpret_cast<float4*>(s.MINV)[q] = make_float4(mi[0], mi[1], mi[2], mi[3]);
        } else {
          float mi[4];
          rs_unpack(reinterpret_cast<const float4*>(s.MINV)[q], mi);
#pragma unroll
          for (int i = 0; i < 4; ++i) sghmc_apply(t[i], v[i], mi[i], gr[i], zf[i], a.s);
        }
        const float4 tn = make_float4(t[0], t[1], t[2], t[3]);
        reinterpret_cast<float4*>(s.TH)[q] = tn;
        reinterpret_cast<float4*>(s.V)[q] = make_float4(v[0], v[1], v[2], v[3]);
        if (snap) __stcs(tr4 + q, tn);
      }
    }
  }
  __syncthreads();
  // ---- the state goes back ----
  {
    const int64_t g0 = chain * n4;
    float4 *t4 = reinterpret_cast<float4*>(a.theta) + g0, *v4 = reinterpret_cast<float4*>(a.v) + g0,
           *ta4 = reinterpret_cast<float4*>(a.tau) + g0, *g4 = reinterpret_cast<float4*>(a.g) + g0,
           *h4 = reinterpret_cast<float4*>(a.v_hat) + g0, *m4 = reinterpret_cast<float4*>(a.minv) + g0;
    for (int q = tid; q < n4; q += NT) {
      __stcs(t4 + q, reinterpret_cast<const float4*>(s.TH)[q]);
      __stcs(v4 + q, reinterpret_cast<const float4*>(s.V)[q]);
      __stcs(ta4 + q, reinterpret_cast<const float4*>(s.TAU)[q]);
      __stcs(g4 + q, reinterpret_cast<const float4*>(s.GG)[q]);
      __stcs(h4 + q, reinterpret_cast<const float4*>(s.VH)[q]);
      __stcs(m4 + q, reinterpret_cast<const float4*>(s.MINV)[q]);
    }
  }
}

static int g_resident_threads = RS_T;

template <int NT>
static void launch_resident(const ResidentArgs& a, size_t smem, cudaStream_t st) {
  auto k = bnn_sghmc_resident_kernel<NT>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<(unsigned)a.n_chains, NT, smem, st>>>(a);
}

static bool resident_shape_ok(int n_in, int batch) {
  if (n_in < 1 || n_in > 64 || batch < 1 || batch > RS_MAX_BATCH) return false;
  const BnnLayout L = make_layout(n_in);
  if (L.D % 4 != 0) return false;                   // element groups of 4 must not straddle chains
  const ResidentSmem s = resident_carve(nullptr, batch, n_in, L.D);
  return (size_t)s.total * sizeof(float) <= 227 * 1024;
}

}  // namespace sgmcmc

using namespace sgmcmc;

extern "C" int sgmcmc_set_bnn_resident_threads(int threads) {
  SG_REQUIRE(threads == 0 || threads == 448 || threads == 672 || threads == 1024, SGMCMC_E_INVALID,
             "bnn resident kernel: threads per chain must be 448, 672 or 1024 (0 = default)");
  g_resident_threads = threads == 0 ? RS_T : threads;
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
             "bnn_sghmc_run_resident: n_in = %d (must be odd, <= 64) / batch = %d (<= %d) does not fit an SM's shared "
             "memory or the 4-element update groups; use sgmcmc_bnn_sghmc_run_f32", n_in, batch, RS_MAX_BATCH);
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
  SG_REQUIRE((chain_offset * (uint64_t)a.L.D) % 4 == 0, SGMCMC_E_INVALID, "chain_offset * D must be a multiple of 4");
  a.group_offset = chain_offset * (uint64_t)a.L.D / 4;
  for (float* p : {theta, v, tau, g, v_hat, minv})
    SG_REQUIRE(aligned_to(p, 16), SGMCMC_E_ALIGN, "bnn_sghmc_run_resident: state arrays must be 16-byte aligned");
  SG_REQUIRE((!trace || aligned_to(trace, 16)) && (!z || aligned_to(z, 16)) && (!grad_out || aligned_to(grad_out, 16)),
             SGMCMC_E_ALIGN, "bnn_sghmc_run_resident: trace, z and grad_out must be 16-byte aligned");
  if (n_chains == 0 || n_steps == 0) return SGMCMC_OK;
  SG_REQUIRE(n_chains <= 0x7fffffff, SGMCMC_E_UNSUPPORTED, "bnn_sghmc_run_resident: at most 2^31 - 1 chains per launch");
  const ResidentSmem sm = resident_carve(nullptr, batch, n_in, a.L.D);
  const size_t smem = (size_t)sm.total * sizeof(float);
  switch (g_resident_threads) {
    case 448: launch_resident<448>(a, smem, (cudaStream_t)stream); break;
    case 1024: launch_resident<1024>(a, smem, (cudaStream_t)stream); break;
    default: launch_resident<RS_T>(a, smem, (cudaStream_t)stream); break;
  }
  return check_launch("bnn_sghmc_resident_kernel");
}
