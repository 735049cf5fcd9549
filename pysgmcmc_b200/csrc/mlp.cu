// K4 for ANY fully connected architecture: cost + gradient of a BNN whose network is
// n_in -> h_1 -> ... -> h_L -> 1 (tanh hidden layers, linear head, learned log-variance), i.e.
// pysgmcmc/models/bayesian_neural_network.py:337-388 over a user-chosen `get_net` of the shape
// of get_default_net (:28-69) with other widths / depths -- BASELINE.json configs[4] names the
// wide 1000-512-512 network (D = 777 682).  The 50-50-50 default keeps its specialised kernels
// (bnn_mma.cuh / bnn.cu, a whole chain per CTA in shared memory); here a chain's weights do not
// fit an SM (3.1 MB), so the step is a sequence of batched skinny GEMMs over ALL chains,
// weights streamed from HBM exactly once per pass:
//
//   forward   l = 1..L : H_l = tanh(H_{l-1} W_l + b_l)              mlp_fwd_kernel
//   head               : f, cost, dW_{L+1}, db_{L+1}, drho, dZ_L     mlp_head_kernel
//   backward  l = L..1 : dW_l = H_{l-1}^T dZ_l (+ prior), db_l       mlp_wgrad_kernel
//                        dZ_{l-1} = (dZ_l W_l^T) * (1 - H_{l-1}^2)   mlp_bwd_data_kernel  (l > 1)
//
// Roofline: the minibatch has 20 rows, so every weight is used for 20 FMAs per pass (forward,
// backward-data, weight gradient): 120 flop per 4-byte weight read + 4-byte gradient written =
// 15 flop/B -- HBM bound on paper (6.2 MB of theta + grad per chain-step at D = 777 682), FP32
// pipe bound in practice once the operand loads from shared memory are counted, which is why
// each thread owns 4 columns (or 4 rows) x all batch rows: one 128-bit weight load and BT/4
// 128-bit broadcast loads of the activations feed 4 * BT FFMAs.
// Activations / dZ live in a caller-provided workspace ([C] x mlp_workspace_floats), a few
// hundred KB per chain, L2 resident between the kernels of a step.
//
// Layout of a chain's parameters (the reference's tf.trainable_variables() order, kernels
// [in, out] row-major): W_1 b_1 W_2 b_2 ... W_{L+1} b_{L+1} rho.
#include "bnn_common.cuh"

namespace sgmcmc {

constexpr int MLP_MAX_W = 8;        // weight matrices: up to 7 hidden layers + the head
constexpr int MLP_THREADS = 128;
constexpr int MLP_KC = 64;          // rows of the staged operand per shared-memory chunk

struct MlpLayout {
  int n_w;                          // weight matrices = hidden layers + 1
  int width[MLP_MAX_W + 1];         // width[0] = n_in, width[1..n_w-1] hidden, width[n_w] = 1
  int64_t oW[MLP_MAX_W], ob[MLP_MAX_W], orho, D;
  int wp[MLP_MAX_W + 1];            // widths rounded up to 4 (row strides in the workspace)
  int64_t oH[MLP_MAX_W], oZ[MLP_MAX_W];   // workspace offsets (floats) of H_l and dZ_l, l = 1..n_w-1
  int64_t oSq, oDf, ws_floats;      // partial sums of squares (prior), d cost / d f, total per chain
  int sq_slot[MLP_MAX_W];           // first partial-sum slot of layer l (one slot per column tile)
  int n_sq;
};

struct MlpArgs {
  const float* theta;      // [n_theta_rows, D]
  const float* X;          // [n_rows, n_in]
  const float* y;          // [n_rows] (NULL for predict)
  const int32_t* starts;   // [C] first row of each chain's minibatch, or NULL (see row0)
  float* ws;               // [C, ws_floats]
  float* cost;             // [C] or NULL
  float* grad;             // [C, D] or NULL
  float* mse;              // [C] or NULL
  float* fout;             // predict: [n_theta_rows, n_rows, 2] or NULL
  int64_t n_chains;        // work items: chains (training) or nets x row tiles (predict)
  int64_t n_rows;          // rows of X
  int theta_div;           // theta row of work item c = c / theta_div; its row tile = c % theta_div
  int batch;               // rows per work item
  float inv_bs, inv_n, prior_den_inv;
  MlpLayout L;
};

static int make_mlp_layout(MlpLayout& L, const int* widths, int n_widths, int batch) {
  SG_REQUIRE(widths != nullptr && n_widths >= 3 && n_widths <= MLP_MAX_W + 1, SGMCMC_E_UNSUPPORTED,
             "mlp: widths = [n_in, hidden..., 1] with 1 to %d hidden layers (got %d entries)", MLP_MAX_W - 1,
             n_widths);
  SG_REQUIRE(widths[n_widths - 1] == 1, SGMCMC_E_UNSUPPORTED, "mlp: the output width must be 1");
  L.n_w = n_widths - 1;
  int64_t o = 0, w = 0;
  int slot = 0;
  for (int l = 0; l <= L.n_w; ++l) {
    SG_REQUIRE(widths[l] >= 1 && widths[l] <= (1 << 20), SGMCMC_E_UNSUPPORTED, "mlp: layer width out of range");
    L.width[l] = widths[l];
    L.wp[l] = (widths[l] + 3) & ~3;
  }
  for (int l = 0; l < L.n_w; ++l) {
    L.oW[l] = o; o += (int64_t)L.width[l] * L.width[l + 1];
    L.ob[l] = o; o += L.width[l + 1];
    L.sq_slot[l] = slot;
    slot += (L.width[l + 1] + MLP_THREADS - 1) / MLP_THREADS;     // >= the column tiles of either CPT
  }
  L.orho = o; o += 1;
  L.D = o;
  L.n_sq = slot;
  const int bp = (batch + 3) & ~3;
  for (int l = 1; l < L.n_w; ++l) {
    L.oH[l] = w; w += (int64_t)bp * L.wp[l];
    L.oZ[l] = w; w += (int64_t)bp * L.wp[l];
  }
  L.oH[0] = L.oZ[0] = 0;
  L.oSq = w; w += (slot + 3) & ~3;
  L.oDf = w; w += 32;
  L.ws_floats = (w + 3) & ~3;
  return SGMCMC_OK;
}

// columns per thread of the column-owner kernels (forward, weight gradient) for a layer of
// n_out units, and the number of column tiles (CTAs per chain) that gives
__host__ __device__ __forceinline__ int mlp_cpt(int n_out) { return n_out >= 256 ? 4 : 1; }
__host__ __device__ __forceinline__ int mlp_col_tiles(int n_out) {
  return (n_out + mlp_cpt(n_out) * MLP_THREADS - 1) / (mlp_cpt(n_out) * MLP_THREADS);
}

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float s = 0.0f;
  for (int i = 0; i < nw; ++i) s += red[i];
  return s;
}

// first row of work item `chain` and how many of its `batch` rows exist
__device__ __forceinline__ void item_rows(const MlpArgs& a, int64_t chain, int64_t& row0, int& rows) {
  row0 = a.starts != nullptr ? (int64_t)a.starts[chain] : (chain % a.theta_div) * (int64_t)a.batch;
  const int64_t left = a.n_rows - row0;
  rows = (int)(left < a.batch ? (left < 0 ? 0 : left) : a.batch);
}

// Stage rows [k0, k0 + kc) of the left operand of layer l (H_{l-1}, or the minibatch for l = 1)
// transposed into shared memory: sH[ii * BT + b] = H_{l-1}[b][k0 + ii]; rows b >= rows are 0.
template <int BT>
__device__ __forceinline__ void stage_left(const MlpArgs& a, int l, const float* __restrict__ ws, int64_t row0,
                                           int rows, int k0, int kc, float* __restrict__ sH) {
  const int n_in = a.L.width[l - 1];
  for (int e = threadIdx.x; e < kc * BT; e += blockDim.x) {
    const int b = e / kc, ii = e - b * kc;                       // ii fastest: coalesced global reads
    float v = 0.0f;
    if (b < rows)
      v = l == 1 ? __ldg(a.X + (row0 + b) * n_in + k0 + ii) : ws[a.L.oH[l - 1] + (int64_t)b * a.L.wp[l - 1] + k0 + ii];
    sH[ii * BT + b] = v;
  }
}

// ---- forward: H_l = tanh(H_{l-1} W_l + b_l), thread = CPT consecutive columns x all rows ----------
template <int BT, int CPT>
__global__ void __launch_bounds__(MLP_THREADS) mlp_fwd_kernel(MlpArgs a, int l) {
  __shared__ __align__(16) float sH[MLP_KC * BT];
  __shared__ float red[MLP_THREADS / 32];
  const int n_in = a.L.width[l - 1], n_out = a.L.width[l];
  const int n_ct = (n_out + CPT * MLP_THREADS - 1) / (CPT * MLP_THREADS);
  const int64_t chain = blockIdx.x / n_ct;
  const int ct = (int)(blockIdx.x - chain * n_ct);
  const float* __restrict__ th = a.theta + (chain / a.theta_div) * a.L.D;
  const float* __restrict__ W = th + a.L.oW[l - 1];
  const float* __restrict__ bias = th + a.L.ob[l - 1];
  float* __restrict__ ws = a.ws + chain * a.L.ws_floats;
  int64_t row0;
  int rows;
  item_rows(a, chain, row0, rows);
  const int j0 = (ct * MLP_THREADS + (int)threadIdx.x) * CPT;
  const bool vec = CPT == 4 && (n_out & 3) == 0 && aligned_to_dev(W, 16);
  const bool vec2 = CPT == 4 && (n_out & 1) == 0 && aligned_to_dev(W, 8);   // (rows of odd chains when D % 4 == 2)
  float acc[BT][CPT];
#pragma unroll
  for (int b = 0; b < BT; ++b)
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[b][c] = 0.0f;
  float sq = 0.0f;
  for (int k0 = 0; k0 < n_in; k0 += MLP_KC) {
    const int kc = min(MLP_KC, n_in - k0);
    __syncthreads();
    stage_left<BT>(a, l, ws, row0, rows, k0, kc, sH);
    __syncthreads();
    if (j0 < n_out) {
      const float* __restrict__ wrow = W + (int64_t)k0 * n_out + j0;
#pragma unroll 4
      for (int ii = 0; ii < kc; ++ii) {
        float w[CPT];
        if (vec) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(wrow + (int64_t)ii * n_out));
          w[0] = q.x;
          if constexpr (CPT == 4) { w[1] = q.y; w[2] = q.z; w[3] = q.w; }
        } else if (vec2 && j0 + 3 < n_out) {
          const float2* p2 = reinterpret_cast<const float2*>(wrow + (int64_t)ii * n_out);
          const float2 q0 = __ldg(p2), q1 = __ldg(p2 + 1);
          w[0] = q0.x;
          if constexpr (CPT == 4) { w[1] = q0.y; w[2] = q1.x; w[3] = q1.y; }
        } else {
#pragma unroll
          for (int c = 0; c < CPT; ++c) w[c] = j0 + c < n_out ? __ldg(wrow + (int64_t)ii * n_out + c) : 0.0f;
        }
#pragma unroll
        for (int c = 0; c < CPT; ++c) sq = fmaf(w[c], w[c], sq);
        const float4* h4 = reinterpret_cast<const float4*>(sH + ii * BT);
#pragma unroll
        for (int b4 = 0; b4 < BT / 4; ++b4) {
          const float4 h = h4[b4];
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            acc[4 * b4 + 0][c] = fmaf(h.x, w[c], acc[4 * b4 + 0][c]);
            acc[4 * b4 + 1][c] = fmaf(h.y, w[c], acc[4 * b4 + 1][c]);
            acc[4 * b4 + 2][c] = fmaf(h.z, w[c], acc[4 * b4 + 2][c]);
            acc[4 * b4 + 3][c] = fmaf(h.w, w[c], acc[4 * b4 + 3][c]);
          }
        }
      }
    }
  }
  if (j0 < n_out) {
    float* __restrict__ Hout = ws + a.L.oH[l];
    const int wp = a.L.wp[l];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      if (j0 + c < n_out) {
        const float bj = __ldg(bias + j0 + c);
        sq = fmaf(bj, bj, sq);
#pragma unroll
        for (int b = 0; b < BT; ++b)
          if (b < rows) Hout[(int64_t)b * wp + j0 + c] = fast_tanh(acc[b][c] + bj);
      }
    }
  }
  // sum of squares of this tile's weights and biases, for the weight prior (:131-141)
  const float tot = block_sum(sq, red);
  if (threadIdx.x == 0) ws[a.L.oSq + a.L.sq_slot[l - 1] + ct] = tot;
}

// ---- head: f = H_L W_{L+1} + b_{L+1}; loss (bayesian_neural_network.py:368-388); gradient of the
// head parameters and rho; dZ_L = (df W_{L+1}^T) * (1 - H_L^2).  One CTA per work item. -------------
template <bool WANT_GRAD>
__global__ void __launch_bounds__(MLP_THREADS) mlp_head_kernel(MlpArgs a) {
  __shared__ float sF[32], sDf[32], red[MLP_THREADS / 32];
  const int l = a.L.n_w;                       // the head is weight matrix n_w (1-based)
  const int hL = a.L.width[l - 1], wp = a.L.wp[l - 1];
  const int64_t chain = blockIdx.x;
  const int64_t trow = chain / a.theta_div;
  const float* __restrict__ th = a.theta + trow * a.L.D;
  const float* __restrict__ W = th + a.L.oW[l - 1];
  float* __restrict__ ws = a.ws + chain * a.L.ws_floats;
  const float* __restrict__ H = ws + a.L.oH[l - 1];
  int64_t row0;
  int rows;
  item_rows(a, chain, row0, rows);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float b4 = th[a.L.ob[l - 1]];
  const float rho = th[a.L.orho];
  for (int b = w; b < rows; b += MLP_THREADS / 32) {
    float f = 0.0f;
    for (int i = lane; i < hL; i += 32) f = fmaf(H[(int64_t)b * wp + i], __ldg(W + i), f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) f += __shfl_xor_sync(0xffffffffu, f, o);
    if (lane == 0) sF[b] = f + b4;
  }
  __syncthreads();
  if (a.fout != nullptr) {                     // predict: (mean, log variance) per row (:535-557)
    for (int b = tid; b < rows; b += MLP_THREADS) {
      float* o = a.fout + (trow * a.n_rows + row0 + b) * 2;
      o[0] = sF[b];
      o[1] = rho;
    }
  }
  if (a.y == nullptr) return;
  const float e_rho = expf(rho);
  const float fvi = 1.0f / (e_rho + 1e-16f);                          // :368
  float sse = 0.0f;
  if (tid < 32) {
    float diff = tid < rows ? __ldg(a.y + row0 + tid) - sF[tid] : 0.0f;
    if (tid < rows) sDf[tid] = -(diff * fvi) * a.inv_bs;              // d cost / d f_i
    sse = diff * diff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sse += __shfl_xor_sync(0xffffffffu, sse, o);
  }
  // squares of the head's own parameters
  float sq = 0.0f;
  for (int i = tid; i < hL; i += MLP_THREADS) { const float v = __ldg(W + i); sq = fmaf(v, v, sq); }
  const float sq_head = block_sum(sq, red);                           // (also orders sDf / sse)
  const float pscale = a.prior_den_inv * a.inv_n;
  if (tid == 0) {
    float sq_t = sq_head + b4 * b4 + rho * rho;
    for (int q = 1; q < l; ++q)                                       // one partial per column tile of layer q
      for (int ct = 0; ct < mlp_col_tiles(a.L.width[q]); ++ct) sq_t += ws[a.L.oSq + a.L.sq_slot[q - 1] + ct];
    const float lv_den = 0.02f + 3e-16f;                              // safe_divide(., 2 * var)
    const float dl = rho - logf(1e-6f);
    const float log_like_data = (-sse * (0.5f * fvi) - 0.5f * rho * (float)rows) * a.inv_bs;
    const float lv = -(dl * dl) / lv_den - 0.5f * logf(0.01f);        // :102-107
    const float wp_ = (-0.5f * sq_t) * a.prior_den_inv;               // :131-141
    if (a.cost != nullptr) a.cost[chain] = -(log_like_data + (lv + wp_) * a.inv_n);
    if (a.mse != nullptr) a.mse[chain] = sse / (float)rows;
    if (WANT_GRAD) {
      float* g = a.grad + chain * a.L.D;
      const float drho_data = -(0.5f * sse * e_rho * fvi * fvi - 0.5f * (float)rows) * a.inv_bs;
      g[a.L.orho] = drho_data + (2.0f * dl / lv_den) * a.inv_n + rho * pscale;
      float db = 0.0f;
      for (int b = 0; b < rows; ++b) db += sDf[b];
      g[a.L.ob[l - 1]] = fmaf(b4, pscale, db);
    }
  }
  if (!WANT_GRAD) return;
  float* __restrict__ g = a.grad + chain * a.L.D + a.L.oW[l - 1];
  float* __restrict__ dZ = ws + a.L.oZ[l - 1];
  for (int i = tid; i < hL; i += MLP_THREADS) {
    const float wi = __ldg(W + i);
    float dw = 0.0f;
    for (int b = 0; b < rows; ++b) {
      const float h = H[(int64_t)b * wp + i];
      dw = fmaf(h, sDf[b], dw);
      dZ[(int64_t)b * wp + i] = (sDf[b] * wi) * fmaf(-h, h, 1.0f);
    }
    g[i] = fmaf(wi, pscale, dw);
  }
}

// ---- weight gradient: dW_l[i][j] = sum_b H_{l-1}[b][i] dZ_l[b][j] + pscale W_l[i][j] (and db_l);
// thread = CPT consecutive columns (dZ of those columns in registers), the rows i of its slice ------
template <int BT, int CPT>
__global__ void __launch_bounds__(MLP_THREADS) mlp_wgrad_kernel(MlpArgs a, int l, int n_is) {
  __shared__ __align__(16) float sH[MLP_KC * BT];
  const int n_in = a.L.width[l - 1], n_out = a.L.width[l];
  const int n_ct = (n_out + CPT * MLP_THREADS - 1) / (CPT * MLP_THREADS);
  const int64_t chain = blockIdx.x / (n_ct * n_is);
  const int rem = (int)(blockIdx.x - chain * (n_ct * n_is));
  const int is = rem / n_ct, ct = rem - is * n_ct;
  const float* __restrict__ th = a.theta + (chain / a.theta_div) * a.L.D;
  const float* __restrict__ W = th + a.L.oW[l - 1];
  float* __restrict__ gW = a.grad + chain * a.L.D + a.L.oW[l - 1];
  float* __restrict__ ws = a.ws + chain * a.L.ws_floats;
  int64_t row0;
  int rows;
  item_rows(a, chain, row0, rows);
  const int j0 = (ct * MLP_THREADS + (int)threadIdx.x) * CPT;
  const float pscale = a.prior_den_inv * a.inv_n;
  const bool vec = CPT == 4 && (n_out & 3) == 0 && aligned_to_dev(W, 16) && aligned_to_dev(gW, 16);
  const bool vec2 = CPT == 4 && (n_out & 1) == 0 && aligned_to_dev(W, 8) && aligned_to_dev(gW, 8);
  float dz[BT][CPT];
  const float* __restrict__ dZ = ws + a.L.oZ[l];
  const int wp = a.L.wp[l];
#pragma unroll
  for (int b = 0; b < BT; ++b)
#pragma unroll
    for (int c = 0; c < CPT; ++c) dz[b][c] = (b < rows && j0 + c < n_out) ? dZ[(int64_t)b * wp + j0 + c] : 0.0f;
  if (is == 0 && j0 < n_out) {
#pragma unroll
    for (int c = 0; c < CPT; ++c)
      if (j0 + c < n_out) {
        float db = 0.0f;
#pragma unroll
        for (int b = 0; b < BT; ++b) db += dz[b][c];
        a.grad[chain * a.L.D + a.L.ob[l - 1] + j0 + c] = fmaf(__ldg(th + a.L.ob[l - 1] + j0 + c), pscale, db);
      }
  }
  const int ilen = (n_in + n_is - 1) / n_is;
  const int i_begin = is * ilen, i_end = min(n_in, i_begin + ilen);
  for (int k0 = i_begin; k0 < i_end; k0 += MLP_KC) {
    const int kc = min(MLP_KC, i_end - k0);
    __syncthreads();
    stage_left<BT>(a, l, ws, row0, rows, k0, kc, sH);
    __syncthreads();
    if (j0 < n_out) {
#pragma unroll 2
      for (int ii = 0; ii < kc; ++ii) {
        const int64_t o = (int64_t)(k0 + ii) * n_out + j0;
        float w[CPT], gsum[CPT];
        if (vec) {
          const float4 q = __ldg(reinterpret_cast<const float4*>(W + o));
          w[0] = q.x;
          if constexpr (CPT == 4) { w[1] = q.y; w[2] = q.z; w[3] = q.w; }
        } else if (vec2 && j0 + 3 < n_out) {
          const float2 q0 = __ldg(reinterpret_cast<const float2*>(W + o)), q1 = __ldg(reinterpret_cast<const float2*>(W + o) + 1);
          w[0] = q0.x;
          if constexpr (CPT == 4) { w[1] = q0.y; w[2] = q1.x; w[3] = q1.y; }
        } else {
#pragma unroll
          for (int c = 0; c < CPT; ++c) w[c] = j0 + c < n_out ? __ldg(W + o + c) : 0.0f;
        }
#pragma unroll
        for (int c = 0; c < CPT; ++c) gsum[c] = 0.0f;
        const float4* h4 = reinterpret_cast<const float4*>(sH + ii * BT);
#pragma unroll
        for (int b4 = 0; b4 < BT / 4; ++b4) {
          const float4 h = h4[b4];
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            gsum[c] = fmaf(h.x, dz[4 * b4 + 0][c], gsum[c]);
            gsum[c] = fmaf(h.y, dz[4 * b4 + 1][c], gsum[c]);
            gsum[c] = fmaf(h.z, dz[4 * b4 + 2][c], gsum[c]);
            gsum[c] = fmaf(h.w, dz[4 * b4 + 3][c], gsum[c]);
          }
        }
        if (vec) {
          if constexpr (CPT == 4)
            *reinterpret_cast<float4*>(gW + o) = make_float4(fmaf(w[0], pscale, gsum[0]), fmaf(w[1], pscale, gsum[1]),
                                                             fmaf(w[2], pscale, gsum[2]), fmaf(w[3], pscale, gsum[3]));
        } else if (vec2 && j0 + 3 < n_out) {
          if constexpr (CPT == 4) {
            float2* g2 = reinterpret_cast<float2*>(gW + o);
            g2[0] = make_float2(fmaf(w[0], pscale, gsum[0]), fmaf(w[1], pscale, gsum[1]));
            g2[1] = make_float2(fmaf(w[2], pscale, gsum[2]), fmaf(w[3], pscale, gsum[3]));
          }
        } else {
#pragma unroll
          for (int c = 0; c < CPT; ++c)
            if (j0 + c < n_out) gW[o + c] = fmaf(w[c], pscale, gsum[c]);
        }
      }
    }
  }
}

// ---- backward data: dZ_{l-1}[b][i] = (sum_j dZ_l[b][j] W_l[i][j]) * (1 - H_{l-1}[b][i]^2);
// thread = RI rows i (interleaved, so a warp's rows are consecutive) x all batch rows -----------------
template <int BT, int RI>
__global__ void __launch_bounds__(MLP_THREADS) mlp_bwd_data_kernel(MlpArgs a, int l) {
  __shared__ __align__(16) float sZ[MLP_KC * BT];
  const int n_in = a.L.width[l - 1], n_out = a.L.width[l];
  const int n_rt = (n_in + RI * MLP_THREADS - 1) / (RI * MLP_THREADS);
  const int64_t chain = blockIdx.x / n_rt;
  const int rt = (int)(blockIdx.x - chain * n_rt);
  const float* __restrict__ th = a.theta + (chain / a.theta_div) * a.L.D;
  const float* __restrict__ W = th + a.L.oW[l - 1];
  float* __restrict__ ws = a.ws + chain * a.L.ws_floats;
  int64_t row0;
  int rows;
  item_rows(a, chain, row0, rows);
  const float* __restrict__ dZ = ws + a.L.oZ[l];
  const int wpz = a.L.wp[l];
  int irow[RI];
#pragma unroll
  for (int r = 0; r < RI; ++r) irow[r] = (rt * RI + r) * MLP_THREADS + (int)threadIdx.x;
  // rows of W_l start at multiples of n_out floats: 128-bit loads need n_out % 4 == 0 and an
  // aligned base; a base that is only 8-byte aligned (odd chains when D % 4 == 2) takes two 64-bit loads
  const bool vec = (n_out & 3) == 0 && aligned_to_dev(W, 8);
  const bool a16 = aligned_to_dev(W, 16);
  float acc[RI][BT];
#pragma unroll
  for (int r = 0; r < RI; ++r)
#pragma unroll
    for (int b = 0; b < BT; ++b) acc[r][b] = 0.0f;
  for (int k0 = 0; k0 < n_out; k0 += MLP_KC) {
    const int kc = min(MLP_KC, n_out - k0);
    __syncthreads();
    for (int e = threadIdx.x; e < kc * BT; e += blockDim.x) {      // sZ[jj * BT + b] = dZ_l[b][k0 + jj]
      const int b = e / kc, jj = e - b * kc;
      sZ[jj * BT + b] = b < rows ? dZ[(int64_t)b * wpz + k0 + jj] : 0.0f;
    }
    __syncthreads();
    if (vec && (kc & 3) == 0) {
      for (int jj = 0; jj < kc; jj += 4) {
        float w[RI][4];
#pragma unroll
        for (int r = 0; r < RI; ++r) {
          if (irow[r] < n_in) {
            const float* wp4 = W + (int64_t)irow[r] * n_out + k0 + jj;
            if (a16) {
              const float4 q = __ldg(reinterpret_cast<const float4*>(wp4));
              w[r][0] = q.x; w[r][1] = q.y; w[r][2] = q.z; w[r][3] = q.w;
            } else {
              const float2 q0 = __ldg(reinterpret_cast<const float2*>(wp4)), q1 = __ldg(reinterpret_cast<const float2*>(wp4) + 1);
              w[r][0] = q0.x; w[r][1] = q0.y; w[r][2] = q1.x; w[r][3] = q1.y;
            }
          } else {
            w[r][0] = w[r][1] = w[r][2] = w[r][3] = 0.0f;
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4* z4 = reinterpret_cast<const float4*>(sZ + (jj + c) * BT);
#pragma unroll
          for (int b4 = 0; b4 < BT / 4; ++b4) {
            const float4 z = z4[b4];
#pragma unroll
            for (int r = 0; r < RI; ++r) {
              acc[r][4 * b4 + 0] = fmaf(z.x, w[r][c], acc[r][4 * b4 + 0]);
              acc[r][4 * b4 + 1] = fmaf(z.y, w[r][c], acc[r][4 * b4 + 1]);
              acc[r][4 * b4 + 2] = fmaf(z.z, w[r][c], acc[r][4 * b4 + 2]);
              acc[r][4 * b4 + 3] = fmaf(z.w, w[r][c], acc[r][4 * b4 + 3]);
            }
          }
        }
      }
    } else {
#pragma unroll 2
      for (int jj = 0; jj < kc; ++jj) {
        float w[RI];
#pragma unroll
        for (int r = 0; r < RI; ++r) w[r] = irow[r] < n_in ? __ldg(W + (int64_t)irow[r] * n_out + k0 + jj) : 0.0f;
        const float4* z4 = reinterpret_cast<const float4*>(sZ + jj * BT);
#pragma unroll
        for (int b4 = 0; b4 < BT / 4; ++b4) {
          const float4 z = z4[b4];
#pragma unroll
          for (int r = 0; r < RI; ++r) {
            acc[r][4 * b4 + 0] = fmaf(z.x, w[r], acc[r][4 * b4 + 0]);
            acc[r][4 * b4 + 1] = fmaf(z.y, w[r], acc[r][4 * b4 + 1]);
            acc[r][4 * b4 + 2] = fmaf(z.z, w[r], acc[r][4 * b4 + 2]);
            acc[r][4 * b4 + 3] = fmaf(z.w, w[r], acc[r][4 * b4 + 3]);
          }
        }
      }
    }
  }
  const float* __restrict__ H = ws + a.L.oH[l - 1];
  float* __restrict__ dZp = ws + a.L.oZ[l - 1];
  const int wp = a.L.wp[l - 1];
#pragma unroll
  for (int r = 0; r < RI; ++r)
    if (irow[r] < n_in) {
#pragma unroll
      for (int b = 0; b < BT; ++b)
        if (b < rows) {
          const float h = H[(int64_t)b * wp + irow[r]];
          dZp[(int64_t)b * wp + irow[r]] = acc[r][b] * fmaf(-h, h, 1.0f);
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------------
template <int BT>
static int launch_mlp_bt(const MlpArgs& a, bool want_grad, cudaStream_t st) {
  const MlpLayout& L = a.L;
  const int n_hidden = L.n_w - 1;
  const int64_t C = a.n_chains;
  for (int l = 1; l <= n_hidden; ++l) {
    const int n_ct = mlp_col_tiles(L.width[l]);
    if (mlp_cpt(L.width[l]) == 4) mlp_fwd_kernel<BT, 4><<<(unsigned)(C * n_ct), MLP_THREADS, 0, st>>>(a, l);
    else mlp_fwd_kernel<BT, 1><<<(unsigned)(C * n_ct), MLP_THREADS, 0, st>>>(a, l);
    if (int rc = check_launch("mlp_fwd_kernel")) return rc;
  }
  if (want_grad) mlp_head_kernel<true><<<(unsigned)C, MLP_THREADS, 0, st>>>(a);
  else mlp_head_kernel<false><<<(unsigned)C, MLP_THREADS, 0, st>>>(a);
  if (int rc = check_launch("mlp_head_kernel")) return rc;
  if (!want_grad) return SGMCMC_OK;
  for (int l = n_hidden; l >= 1; --l) {
    const int n_in = L.width[l - 1], n_out = L.width[l];
    // split the rows of W_l over CTAs until the grid fills the machine (few chains, wide layers)
    const bool quad = mlp_cpt(n_out) == 4;
    const int n_ct = mlp_col_tiles(n_out);
    int n_is = 1;
    while (C * n_ct * n_is < 592 && n_is * 2 * MLP_KC <= n_in) n_is *= 2;
    if (quad) mlp_wgrad_kernel<BT, 4><<<(unsigned)(C * n_ct * n_is), MLP_THREADS, 0, st>>>(a, l, n_is);
    else mlp_wgrad_kernel<BT, 1><<<(unsigned)(C * n_ct * n_is), MLP_THREADS, 0, st>>>(a, l, n_is);
    if (int rc = check_launch("mlp_wgrad_kernel")) return rc;
    if (l > 1) {
      if (n_in >= 4 * MLP_THREADS && C * ((n_in + 4 * MLP_THREADS - 1) / (4 * MLP_THREADS)) >= 296) {
        const int n_rt = (n_in + 4 * MLP_THREADS - 1) / (4 * MLP_THREADS);
        mlp_bwd_data_kernel<BT, 4><<<(unsigned)(C * n_rt), MLP_THREADS, 0, st>>>(a, l);
      } else {
        const int n_rt = (n_in + MLP_THREADS - 1) / MLP_THREADS;
        mlp_bwd_data_kernel<BT, 1><<<(unsigned)(C * n_rt), MLP_THREADS, 0, st>>>(a, l);
      }
      if (int rc = check_launch("mlp_bwd_data_kernel")) return rc;
    }
  }
  return SGMCMC_OK;
}

static int launch_mlp(const MlpArgs& a, bool want_grad, cudaStream_t st) {
  if (a.batch <= 8) return launch_mlp_bt<8>(a, want_grad, st);
  if (a.batch <= 16) return launch_mlp_bt<16>(a, want_grad, st);
  if (a.batch <= 20) return launch_mlp_bt<20>(a, want_grad, st);
  return launch_mlp_bt<32>(a, want_grad, st);
}

}  // namespace sgmcmc

using namespace sgmcmc;

extern "C" int64_t sgmcmc_mlp_n_params(const int* widths, int n_widths) {
  MlpLayout L;
  if (make_mlp_layout(L, widths, n_widths, 32) != SGMCMC_OK) return -1;
  return L.D;
}

extern "C" int64_t sgmcmc_mlp_workspace_bytes(const int* widths, int n_widths, int64_t n_items, int batch) {
  MlpLayout L;
  if (n_items < 0 || batch < 1 || batch > 32 || make_mlp_layout(L, widths, n_widths, batch) != SGMCMC_OK) return -1;
  return (int64_t)sizeof(float) * L.ws_floats * n_items;
}

extern "C" int sgmcmc_mlp_nll_grad_f32(const float* theta, const float* X, const float* y, const int32_t* starts,
                                       float* cost, float* grad, float* mse, void* workspace,
                                       int64_t workspace_bytes, int64_t n_chains, const int* widths, int n_widths,
                                       int batch, float batch_size_cfg, int64_t n_examples, void* stream) {
  SG_REQUIRE(n_chains >= 0, SGMCMC_E_INVALID, "n_chains must be >= 0");
  SG_REQUIRE(theta && X && y && cost && workspace, SGMCMC_E_INVALID,
             "mlp: theta, X, y, cost and workspace must not be NULL");
  SG_REQUIRE(batch >= 1 && batch <= 32, SGMCMC_E_UNSUPPORTED, "mlp: batch must be in [1, 32] (got %d)", batch);
  SG_REQUIRE(batch_size_cfg > 0 && n_examples >= 1, SGMCMC_E_INVALID, "mlp: batch_size_cfg and n_examples must be > 0");
  MlpArgs a;
  if (int rc = make_mlp_layout(a.L, widths, n_widths, batch)) return rc;
  SG_REQUIRE(aligned_to(workspace, 16), SGMCMC_E_ALIGN, "mlp: workspace must be 16-byte aligned");
  SG_REQUIRE(workspace_bytes >= (int64_t)sizeof(float) * a.L.ws_floats * n_chains, SGMCMC_E_INVALID,
             "mlp: workspace too small (sgmcmc_mlp_workspace_bytes)");
  SG_REQUIRE(aligned_to(theta, 4) && (grad == nullptr || aligned_to(grad, 4)), SGMCMC_E_ALIGN, "mlp: misaligned pointer");
  SG_REQUIRE(n_chains * (int64_t)((a.L.width[1] + MLP_THREADS - 1) / MLP_THREADS) * 64 < (int64_t)1 << 31,
             SGMCMC_E_UNSUPPORTED, "mlp: grid too large");
  if (n_chains == 0) return SGMCMC_OK;
  a.theta = theta; a.X = X; a.y = y; a.starts = starts; a.ws = (float*)workspace;
  a.cost = cost; a.grad = grad; a.mse = mse; a.fout = nullptr;
  a.n_chains = n_chains; a.n_rows = n_examples; a.theta_div = 1; a.batch = batch;
  a.inv_bs = 1.0f / batch_size_cfg;
  a.inv_n = 1.0f / (float)n_examples;
  a.prior_den_inv = 1.0f / ((float)a.L.D + 3e-16f);
  return launch_mlp(a, grad != nullptr, (cudaStream_t)stream);
}

// K10 for any architecture: out[k, i, 0] = f(x_i; theta_k), out[k, i, 1] = rho_k
extern "C" int sgmcmc_mlp_predict_f32(const float* theta, const float* X, float* out, void* workspace,
                                      int64_t workspace_bytes, int64_t n_nets, const int* widths, int n_widths,
                                      int64_t n_points, void* stream) {
  SG_REQUIRE(n_nets >= 0 && n_points >= 0, SGMCMC_E_INVALID, "negative size");
  if (n_nets == 0 || n_points == 0) return SGMCMC_OK;
  SG_REQUIRE(theta && X && out && workspace, SGMCMC_E_INVALID, "mlp_predict: NULL pointer");
  MlpArgs a;
  const int batch = 32;
  if (int rc = make_mlp_layout(a.L, widths, n_widths, batch)) return rc;
  const int64_t tiles = (n_points + batch - 1) / batch;
  SG_REQUIRE(tiles <= (1 << 30) / (n_nets > 0 ? n_nets : 1), SGMCMC_E_UNSUPPORTED, "mlp_predict: too many points");
  SG_REQUIRE(aligned_to(workspace, 16), SGMCMC_E_ALIGN, "mlp: workspace must be 16-byte aligned");
  SG_REQUIRE(workspace_bytes >= (int64_t)sizeof(float) * a.L.ws_floats * n_nets * tiles, SGMCMC_E_INVALID,
             "mlp_predict: workspace too small (sgmcmc_mlp_workspace_bytes with n_items = n_nets * ceil(n_points / 32), "
             "batch = 32)");
  a.theta = theta; a.X = X; a.y = nullptr; a.starts = nullptr; a.ws = (float*)workspace;
  a.cost = nullptr; a.grad = nullptr; a.mse = nullptr; a.fout = out;
  a.n_chains = n_nets * tiles; a.n_rows = n_points; a.theta_div = (int)tiles; a.batch = batch;
  a.inv_bs = 1.0f; a.inv_n = 1.0f; a.prior_den_inv = 1.0f;
  return launch_mlp(a, false, (cudaStream_t)stream);
}
