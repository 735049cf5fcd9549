// K4: per-chain fused BNN cost + gradient (the BOHAMIANN network n_in-50-50-50-1, tanh,
// Gaussian likelihood with a learned log-variance; pysgmcmc/models/
// bayesian_neural_network.py:28-69, :77-141, :337-388 and tf.gradients over it),
// K5: the host-side pipeline K4 -> K1 for `n_steps` steps of BNN-SGHMC,
// K10: the forward pass over stored networks for the predictive.
//
// Work decomposition (K4).  A chain's step is 6 small GEMMs over a 20-row minibatch plus
// the 1-wide input / output layers (305 k FFMA).  A CTA owns NC chains; TPC threads
// cooperate on one chain and thread u owns COLS of the 50 hidden units (unit j = u + c*TPC).
// Every GEMM phase is "weights stationary in registers": the thread loads its column
// (forward), row (backward-data) of the 50x50 kernel straight from global memory into 52
// registers per owned unit and streams the activations of the minibatch from shared memory
// with 128-bit warp-broadcast loads.  The activations H1, H2, H3 ([B x 52] each, rows padded
// 50 -> 52 so every k-loop is 13 float4 steps) are the only shared memory (13 KB per chain);
// backward overwrites them in place with dZ3, dZ2, dZ1.  The gradient leaves as coalesced
// register-to-global stores, so HBM traffic is theta in and grad out (42 KB per chain-step).
//
// What bounds it (profiles/, DESIGN.md "K4"): on this SM a broadcast LDS.128 costs ~2.6
// cycles that do NOT overlap FFMA issue (tools/micro/lds_bcast_bench.cu: time ~= FFMA/4 +
// 2.6 * LDS.128), and each loaded activation word feeds only COLS FFMAs per thread.  COLS is
// therefore the lever (fewer operand loads per FFMA), paid for with 52 registers per unit.
#include "bnn_common.cuh"
#include "bnn_mma.cuh"
namespace sgmcmc {

// acc[r][c] += sum_k act[row r][k] * w[c][k]  for ROWS rows of a [B x HS] activation buffer.
// The 128-bit broadcast loads of step k4+1 are issued before the FFMAs of step k4.
template <int COLS, int ROWS>
__device__ __forceinline__ void dot_rows(const float* __restrict__ act, int i0, int n_rows,
                                         const float (&w)[COLS][HS], float (&acc)[ROWS][COLS]) {
  const float4* rp[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r)
    rp[r] = reinterpret_cast<const float4*>(act + (i0 + (r < n_rows ? r : 0)) * HS);
  float4 hc[ROWS], hn[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) hc[r] = rp[r][0];
#pragma unroll
  for (int k4 = 0; k4 < K4S; ++k4) {
    if (k4 + 1 < K4S) {
#pragma unroll
      for (int r = 0; r < ROWS; ++r) hn[r] = rp[r][k4 + 1];
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int c = 0; c < COLS; ++c) {
        acc[r][c] = fmaf(hc[r].x, w[c][4 * k4 + 0], acc[r][c]);
        acc[r][c] = fmaf(hc[r].y, w[c][4 * k4 + 1], acc[r][c]);
        acc[r][c] = fmaf(hc[r].z, w[c][4 * k4 + 2], acc[r][c]);
        acc[r][c] = fmaf(hc[r].w, w[c][4 * k4 + 3], acc[r][c]);
      }
    if (k4 + 1 < K4S) {
#pragma unroll
      for (int r = 0; r < ROWS; ++r) hc[r] = hn[r];
    }
  }
}

// One dense tanh layer, forward: out[i][j] = tanh(b[j] + sum_k in[i][k] W[k][j]) for the
// thread's columns j; returns sum of squares of the weights it touched (weight prior).
template <int COLS, int TPC, int ROWS>
__device__ __forceinline__ float layer_forward(const float* __restrict__ th, int oW, int ob,
                                               const float* __restrict__ in, float* __restrict__ out,
                                               int batch, int u) {
  float w[COLS][HS], b[COLS];
  float sq = 0.0f;
#pragma unroll
  for (int c = 0; c < COLS; ++c) {
    const int j = u + c * TPC;
    const bool ok = j < HID;          // TPC * COLS may exceed 50 (idle unit slots)
#pragma unroll
    for (int k = 0; k < HID; ++k) {
      w[c][k] = ok ? __ldg(th + oW + k * HID + j) : 0.0f;
      sq = fmaf(w[c][k], w[c][k], sq);
    }
    w[c][HID] = w[c][HID + 1] = 0.0f;
    b[c] = ok ? __ldg(th + ob + j) : 0.0f;
    sq = fmaf(b[c], b[c], sq);
  }
  for (int i0 = 0; i0 < batch; i0 += ROWS) {
    const int n_rows = min(ROWS, batch - i0);
    float acc[ROWS][COLS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int c = 0; c < COLS; ++c) acc[r][c] = b[c];
    dot_rows<COLS, ROWS>(in, i0, n_rows, w, acc);
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
      if (r < n_rows) {
#pragma unroll
        for (int c = 0; c < COLS; ++c)
          if (u + c * TPC < HID) out[(i0 + r) * HS + u + c * TPC] = fast_tanh(acc[r][c]);
      }
  }
  return sq;
}

// In place: h[i][j] <- (sum_m dz[i][m] W[j][m]) * (1 - h[i][j]^2)  for the thread's units j,
// i.e. the activations of the layer below become its dZ.
template <int COLS, int TPC, int ROWS>
__device__ __forceinline__ void layer_backward_data(const float* __restrict__ th, int oW,
                                                    const float* __restrict__ dz, float* __restrict__ h,
                                                    int batch, int u) {
  float w[COLS][HS];
#pragma unroll
  for (int c = 0; c < COLS; ++c) {
    const bool ok = u + c * TPC < HID;
    const float2* row = reinterpret_cast<const float2*>(th + oW + (ok ? u + c * TPC : 0) * HID);
#pragma unroll
    for (int m = 0; m < HID / 2; ++m) {
      const float2 v = ok ? __ldg(row + m) : make_float2(0.0f, 0.0f);
      w[c][2 * m] = v.x;
      w[c][2 * m + 1] = v.y;
    }
    w[c][HID] = w[c][HID + 1] = 0.0f;
  }
  for (int i0 = 0; i0 < batch; i0 += ROWS) {
    const int n_rows = min(ROWS, batch - i0);
    float acc[ROWS][COLS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
      for (int c = 0; c < COLS; ++c) acc[r][c] = 0.0f;
    dot_rows<COLS, ROWS>(dz, i0, n_rows, w, acc);
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
      if (r < n_rows) {
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
          if (u + c * TPC < HID) {
            const int idx = (i0 + r) * HS + u + c * TPC;
            const float hv = h[idx];
            h[idx] = acc[r][c] * fmaf(-hv, hv, 1.0f);
          }
        }
      }
  }
}

// dW[k][j] = sum_i H_prev[i][k] dZ[i][j], db[j] = sum_i dZ[i][j]  (+ weight-prior term),
// written to grad for the thread's columns j.  The k range is walked in KPASS chunks so
// that the COLS x chunk accumulators fit the register budget; every chunk streams only its
// own part of the H_prev rows, so the operand traffic does not grow with KPASS.
template <int COLS, int TPC, int KPASS>
__device__ __forceinline__ void layer_backward_weights(const float* __restrict__ th,
                                                       float* __restrict__ gr, int oW, int ob,
                                                       const float* __restrict__ h_prev,
                                                       const float* __restrict__ dz, int batch, int u,
                                                       float pscale) {
  constexpr int CH = (K4S + KPASS - 1) / KPASS;       // float4 steps per chunk
  const float4* hp = reinterpret_cast<const float4*>(h_prev);
#pragma unroll
  for (int pass = 0; pass < KPASS; ++pass) {
    const int k40 = pass * CH;
    float acc[COLS][4 * CH], db[COLS];
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
      db[c] = 0.0f;
#pragma unroll
      for (int k = 0; k < 4 * CH; ++k) acc[c][k] = 0.0f;
    }
    for (int i = 0; i < batch; ++i) {
      float d[COLS];
#pragma unroll
      for (int c = 0; c < COLS; ++c) {
        d[c] = (u + c * TPC < HID) ? dz[i * HS + u + c * TPC] : 0.0f;
        db[c] += d[c];
      }
      const float4* rp = hp + i * K4S + k40;
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        if (k40 + q < K4S) {
          const float4 h = rp[q];
#pragma unroll
          for (int c = 0; c < COLS; ++c) {
            acc[c][4 * q + 0] = fmaf(h.x, d[c], acc[c][4 * q + 0]);
            acc[c][4 * q + 1] = fmaf(h.y, d[c], acc[c][4 * q + 1]);
            acc[c][4 * q + 2] = fmaf(h.z, d[c], acc[c][4 * q + 2]);
            acc[c][4 * q + 3] = fmaf(h.w, d[c], acc[c][4 * q + 3]);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
      const int j = u + c * TPC;
      if (j < HID) {
#pragma unroll
        for (int kk = 0; kk < 4 * CH; ++kk) {
          const int k = 4 * k40 + kk;
          if (k < HID) gr[oW + k * HID + j] = fmaf(__ldg(th + oW + k * HID + j), pscale, acc[c][kk]);
        }
        if (pass == 0) gr[ob + j] = fmaf(__ldg(th + ob + j), pscale, db[c]);
      }
    }
  }
}

// Launch shape: TPC threads cooperate on one chain, each owning COLS of the 50 hidden units
// (slots with j >= 50 idle), NC chains per CTA.  TPC == 32 is "one warp per chain"
// (__syncwarp only, 64 slots for 50 units); otherwise chains straddle warps and the phases
// are separated by CTA barriers.
template <int COLS, int TPC, int NC>
struct BnnShape {
  static constexpr int THREADS = ((NC * TPC + 31) / 32) * 32;
  static constexpr bool WARP = TPC == 32;
};

template <bool WARP>
__device__ __forceinline__ void chain_sync() {
  if constexpr (WARP) __syncwarp(); else __syncthreads();
}

// shared memory per chain, in floats: X, Y, df, H1, H2, H3, scratch[128]
__host__ __device__ inline int bnn_smem_floats(int batch, int n_in) {
  const int x = ((batch * n_in + 3) / 4) * 4;
  const int yb = ((batch + 3) / 4) * 4;
  return x + 2 * yb + 3 * batch * HS + 128;
}

template <int COLS, int TPC, int NC, int ROWS, int KPASS, int MINB, bool WANT_GRAD>
__global__ void __launch_bounds__(BnnShape<COLS, TPC, NC>::THREADS, MINB)
bnn_nll_grad_kernel(BnnArgs a) {
  constexpr bool WARP = BnnShape<COLS, TPC, NC>::WARP;
  static_assert(TPC <= 64, "scratch layout assumes at most 64 threads per chain");
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int lc = tid / TPC;            // chain slot in this CTA
  const int u = tid - lc * TPC;        // unit group of this thread
  const int batch = a.batch, n_in = a.L.n_in;
  const BnnLayout L = a.L;
  // A CTA walks chain groups blockIdx.x, blockIdx.x + gridDim.x, ...: with gridDim.x ==
  // number of groups this is one group per CTA; a smaller (persistent) grid leaves SM
  // resources free for a concurrently running kernel (sgmcmc_set_bnn_tuning).
  const int64_t n_groups = (a.n_chains + NC - 1) / NC;
  for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
  const int64_t chain = grp * NC + lc;
  const bool active = lc < NC && chain < a.n_chains;

  const int per_chain = bnn_smem_floats(batch, n_in);
  float* sX = smem + (size_t)(lc < NC ? lc : 0) * per_chain;
  float* sY = sX + ((batch * n_in + 3) / 4) * 4;
  float* sDf = sY + ((batch + 3) / 4) * 4;
  float* H1 = sDf + ((batch + 3) / 4) * 4;
  float* H2 = H1 + batch * HS;
  float* H3 = H2 + batch * HS;
  float* sW4 = H3 + batch * HS;        // [64]: the head's weights
  float* scr = sW4 + 64;               // [64]: per-thread partial sums

  const float* th = a.theta + (active ? chain : 0) * L.D;
  float* gr = (WANT_GRAD && a.grad != nullptr) ? a.grad + (active ? chain : 0) * L.D : nullptr;
  float sq = 0.0f;                     // this thread's share of sum(theta^2)

  // ---- P0: stage the minibatch rows X[start : start+B], y[start : start+B] ----
  if (active) {
    const int64_t start = a.starts != nullptr ? a.starts[chain] : 0;
    for (int t = u; t < batch * n_in; t += TPC) sX[t] = __ldg(a.X + start * n_in + t);
    for (int t = u; t < batch; t += TPC) sY[t] = __ldg(a.y + start + t);
    for (int t = u; t < 3 * batch; t += TPC) {     // zero padding columns of H1, H2, H3
      H1[t * HS + HID] = 0.0f;
      H1[t * HS + HID + 1] = 0.0f;
    }
    for (int t = u; t < 64; t += TPC) sW4[t] = t < HID ? __ldg(th + L.oW4 + t) : 0.0f;
  }
  chain_sync<WARP>();

  // ---- P1: layer 1 forward (n_in -> 50) ----
  if (active) {
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
      const int j = u + c * TPC;
      if (j < HID) {
        const float b = __ldg(th + L.ob1 + j);
        sq = fmaf(b, b, sq);
        for (int i = 0; i < batch; ++i) H1[i * HS + j] = b;
        for (int m = 0; m < n_in; ++m) {
          const float w = __ldg(th + L.oW1 + m * HID + j);
          sq = fmaf(w, w, sq);
          for (int i = 0; i < batch; ++i) H1[i * HS + j] = fmaf(sX[i * n_in + m], w, H1[i * HS + j]);
        }
        for (int i = 0; i < batch; ++i) H1[i * HS + j] = fast_tanh(H1[i * HS + j]);
        const float w4 = sW4[j];
        sq = fmaf(w4, w4, sq);
      }
    }
  }
  chain_sync<WARP>();
  // ---- P2, P3: layers 2 and 3 forward ----
  if (active) sq += layer_forward<COLS, TPC, ROWS>(th, L.oW2, L.ob2, H1, H2, batch, u);
  chain_sync<WARP>();
  if (active) sq += layer_forward<COLS, TPC, ROWS>(th, L.oW3, L.ob3, H2, H3, batch, u);
  if (active) scr[u] = sq;
  chain_sync<WARP>();

  // ---- P4: head f[i] = H3[i,:] . W4 + b4, one thread per batch row; loss pieces ----
  if (active) {
    const float b4 = __ldg(th + L.ob4), rho = __ldg(th + L.orho);
    const float fvi = 1.0f / (expf(rho) + 1e-16f);                   // :368
    for (int i = u; i < batch; i += TPC) {
      const float4* hr = reinterpret_cast<const float4*>(H3 + i * HS);
      const float4* wr = reinterpret_cast<const float4*>(sW4);
      float f = b4;
#pragma unroll
      for (int k4 = 0; k4 < K4S; ++k4) {
        const float4 h = hr[k4], w = wr[k4];
        f = fmaf(h.x, w.x, f); f = fmaf(h.y, w.y, f); f = fmaf(h.z, w.z, f); f = fmaf(h.w, w.w, f);
      }
      const float diff = sY[i] - f;
      sDf[i] = -(diff * fvi) * a.inv_bs;                             // d cost / d f_i
      sY[i] = diff * diff;                                           // squared error (:370)
    }
  }
  chain_sync<WARP>();
  if (active && u == 0) {
    // scalar tail of the cost (:372-388) and the rho / b4 gradients, one thread per chain
    const float b4 = __ldg(th + L.ob4), rho = __ldg(th + L.orho);
    const float e_rho = expf(rho);
    const float fvi = 1.0f / (e_rho + 1e-16f);
    float sse = 0.0f, sdf = 0.0f, total_sq = fmaf(b4, b4, rho * rho);
    for (int i = 0; i < batch; ++i) { sse += sY[i]; sdf += sDf[i]; }
    for (int t = 0; t < TPC; ++t) total_sq += scr[t];
    const float log_like_data = (-sse * (0.5f * fvi) - 0.5f * rho * (float)batch) * a.inv_bs;
    const float lv_den = 0.02f + 3e-16f;                             // safe_divide(., 2 * var)
    const float dl = rho - logf(1e-6f);
    const float lv = -(dl * dl) / lv_den - 0.5f * logf(0.01f);       // :102-107
    const float wp = (-0.5f * total_sq) * a.prior_den_inv;           // :131-141
    a.cost[chain] = -(log_like_data + (lv + wp) * a.inv_n);
    if (a.mse != nullptr) a.mse[chain] = sse / (float)batch;
    if (gr != nullptr) {
      const float pscale = a.prior_den_inv * a.inv_n;
      const float drho_data = -(0.5f * sse * e_rho * fvi * fvi - 0.5f * (float)batch) * a.inv_bs;
      gr[L.orho] = drho_data + (2.0f * dl / lv_den) * a.inv_n + rho * pscale;
      gr[L.ob4] = sdf + b4 * pscale;
    }
  }
  if (gr == nullptr) {                 // cost only (uniform across the CTA)
    chain_sync<WARP>();
    continue;
  }

  const float pscale = a.prior_den_inv * a.inv_n;
  // ---- P5: layer 4 backward: dW4, and dZ3 = (df W4^T) * (1 - H3^2) in place over H3 ----
  if (active) {
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
      const int j = u + c * TPC;
      if (j < HID) {
        const float w4 = sW4[j];
        float dw = 0.0f;
        for (int i = 0; i < batch; ++i) {
          const float h = H3[i * HS + j], df = sDf[i];
          dw = fmaf(h, df, dw);
          H3[i * HS + j] = (df * w4) * fmaf(-h, h, 1.0f);
        }
        gr[L.oW4 + j] = fmaf(w4, pscale, dw);
      }
    }
  }
  chain_sync<WARP>();
  // ---- layer 3: dW3 from H2 and dZ3, then H2 <- dZ2 (the rows of H2 must be dead first) ----
  if (active) layer_backward_weights<COLS, TPC, KPASS>(th, gr, L.oW3, L.ob3, H2, H3, batch, u, pscale);
  chain_sync<WARP>();
  if (active) layer_backward_data<COLS, TPC, ROWS>(th, L.oW3, H3, H2, batch, u);
  chain_sync<WARP>();
  // ---- layer 2: dW2 from H1 and dZ2, then H1 <- dZ1 ----
  if (active) layer_backward_weights<COLS, TPC, KPASS>(th, gr, L.oW2, L.ob2, H1, H2, batch, u, pscale);
  chain_sync<WARP>();
  if (active) layer_backward_data<COLS, TPC, ROWS>(th, L.oW2, H2, H1, batch, u);
  // ---- layer 1: dW1 = X^T dZ1, db1 (own column of dZ1 only: no barrier needed) ----
  if (active) {
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
      const int j = u + c * TPC;
      if (j < HID) {
        float db = 0.0f;
        for (int i = 0; i < batch; ++i) db += H1[i * HS + j];
        gr[L.ob1 + j] = fmaf(__ldg(th + L.ob1 + j), pscale, db);
        for (int m = 0; m < n_in; ++m) {
          float dw = 0.0f;
          for (int i = 0; i < batch; ++i) dw = fmaf(sX[i * n_in + m], H1[i * HS + j], dw);
          gr[L.oW1 + m * HID + j] = fmaf(__ldg(th + L.oW1 + m * HID + j), pscale, dw);
        }
      }
    }
  }
  chain_sync<WARP>();                  // shared memory is reused by the next chain group
  }
}

// ---- K10: forward only, one "chain" = (stored network k, block of <= PB test points) ----
constexpr int PB = 32;
template <int COLS, int TPC, int NC>
__global__ void __launch_bounds__(BnnShape<COLS, TPC, NC>::THREADS)
bnn_predict_kernel(const float* __restrict__ theta, const float* __restrict__ X, float* __restrict__ out,
                   int64_t n_nets, int64_t n_points, BnnLayout L) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int lc = tid / TPC, u = tid - lc * TPC;
  const int64_t blocks_per_net = (n_points + PB - 1) / PB;
  const int64_t item = (int64_t)blockIdx.x * NC + lc;
  const bool active = lc < NC && item < n_nets * blocks_per_net;
  const int64_t net = active ? item / blocks_per_net : 0;
  const int64_t p0 = active ? (item % blocks_per_net) * PB : 0;
  const int batch = active ? (int)min((int64_t)PB, n_points - p0) : 0;
  const int n_in = L.n_in;
  const int per_chain = bnn_smem_floats(PB, n_in);
  float* sX = smem + (size_t)(lc < NC ? lc : 0) * per_chain;
  float* H1 = sX + ((PB * n_in + 3) / 4) * 4 + 2 * PB;
  float* H2 = H1 + PB * HS;
  float* H3 = H2 + PB * HS;
  float* sW4 = H3 + PB * HS;
  const float* th = theta + net * L.D;
  if (active) {
    for (int t = u; t < batch * n_in; t += TPC) sX[t] = __ldg(X + p0 * n_in + t);
    for (int t = u; t < 3 * PB; t += TPC) {
      H1[t * HS + HID] = 0.0f;
      H1[t * HS + HID + 1] = 0.0f;
    }
    for (int t = u; t < 64; t += TPC) sW4[t] = t < HID ? __ldg(th + L.oW4 + t) : 0.0f;
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
      const int j = u + c * TPC;
      if (j < HID) {
        const float b = __ldg(th + L.ob1 + j);
        for (int i = 0; i < batch; ++i) {
          float z = b;
          for (int m = 0; m < n_in; ++m) z = fmaf(sX[i * n_in + m], __ldg(th + L.oW1 + m * HID + j), z);
          H1[i * HS + j] = fast_tanh(z);
        }
      }
    }
  }
  __syncthreads();
  if (active) layer_forward<COLS, TPC, 4>(th, L.oW2, L.ob2, H1, H2, batch, u);
  __syncthreads();
  if (active) layer_forward<COLS, TPC, 4>(th, L.oW3, L.ob3, H2, H3, batch, u);
  __syncthreads();
  if (active) {
    const float b4 = __ldg(th + L.ob4), rho = __ldg(th + L.orho);
    for (int i = u; i < batch; i += TPC) {
      float f = b4;
      for (int k = 0; k < HID; ++k) f = fmaf(H3[i * HS + k], sW4[k], f);
      out[(net * n_points + p0 + i) * 2 + 0] = f;
      out[(net * n_points + p0 + i) * 2 + 1] = rho;     // "ones_like(layer_4) * output_bias" (:63-67)
    }
  }
}

// ---- launch shapes (COLS, TPC, NC chains per CTA, ROWS in flight, KPASS, min CTAs / SM),
// selected with sgmcmc_set_bnn_tuning; the default is the fastest of the sweep recorded in
// profiles/ ---------------------------------------------------------------------------
constexpr int K10_COLS = 1, K10_TPC = 50, K10_NC = 5;

// 10-16: tensor-pipe kernel (bnn_mma.cuh) in its accuracy modes; 0-9: FFMA launch shapes.  Default 16: rounded
// hi/lo split, FP32-pipe accumulation of the k-steps' exact hi*hi sums, cross terms in their own accumulator,
// weight fragments split by packed FP32 instructions.  1000-step BNN-SGHMC trajectory at the benchmarked
// shapes: 9.2e-6 from the float32 oracle and 1.32e-5 from the float64 one (the float32 ORACLE: 1.39e-5; mode
// 13: 9.0e-6 / 1.82e-5; mode 10, round 1's default: 4.6e-5 / 4.1e-5; FFMA kernel: 9.8e-6 / 1.5e-5) at 0.238 ms
// for 8192 chains (13: 0.247, 10: 0.202) -- profiles/r02_bnn_trajectory_drift*.jsonl, profiles/r02_k4_variants.jsonl
static int g_bnn_variant = 16;
static int g_bnn_max_ctas = 0;          // 0: one CTA per chain group; > 0: persistent grid of that size
static int64_t g_bnn_chunk = 0;         // chains per K4+K1 chunk inside sgmcmc_bnn_sghmc_run_f32 (0: all)
void set_bnn_chunk(int64_t c) { g_bnn_chunk = c; }
int bnn_variant_count() { return 17; }
void set_bnn_variant(int v) { g_bnn_variant = v; }
void set_bnn_max_ctas(int n) { g_bnn_max_ctas = n; }

template <int COLS, int TPC, int NC, int ROWS, int KPASS, int MINB>
static int launch_variant(const BnnArgs& a, cudaStream_t st) {
  using Shape = BnnShape<COLS, TPC, NC>;
  const size_t smem = (size_t)NC * bnn_smem_floats(a.batch, a.L.n_in) * sizeof(float);
  SG_REQUIRE(smem <= 227 * 1024, SGMCMC_E_UNSUPPORTED,
             "minibatch of %d rows x %d inputs needs %zu B of shared memory per CTA (max 232448)",
             a.batch, a.L.n_in, smem);
  unsigned blocks = (unsigned)((a.n_chains + NC - 1) / NC);
  if (g_bnn_max_ctas > 0 && blocks > (unsigned)g_bnn_max_ctas) blocks = (unsigned)g_bnn_max_ctas;
  if (a.grad != nullptr) {
    auto k = bnn_nll_grad_kernel<COLS, TPC, NC, ROWS, KPASS, MINB, true>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<blocks, Shape::THREADS, smem, st>>>(a);
  } else {
    auto k = bnn_nll_grad_kernel<COLS, TPC, NC, ROWS, KPASS, MINB, false>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<blocks, Shape::THREADS, smem, st>>>(a);
  }
  return check_launch("bnn_nll_grad_kernel");
}

// K4 on the tensor pipe (bnn_mma.cuh): one CTA of ceil(batch / 16) warps per chain.
template <int NB8, int MODE, int MINB = (NB8 > 2 ? 6 : 8), int BATCH_CT = 0>
static int launch_mma(const BnnArgs& a, cudaStream_t st) {
  constexpr int NTHR = 32 * ((NB8 + 1) / 2);
  const size_t smem = (size_t)bnn_mma_smem_floats(a.batch, a.L.n_in, a.L.D) * sizeof(float);
  SG_REQUIRE(smem <= 227 * 1024, SGMCMC_E_UNSUPPORTED, "bnn (mma): %zu B of shared memory per CTA", smem);
  unsigned blocks = (unsigned)a.n_chains;
  if (g_bnn_max_ctas > 0 && blocks > (unsigned)g_bnn_max_ctas) blocks = (unsigned)g_bnn_max_ctas;
  if (a.grad != nullptr) {
    auto k = bnn_mma_kernel<NB8, true, MODE, MINB, BATCH_CT>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    k<<<blocks, NTHR, smem, st>>>(a);
  } else {
    auto k = bnn_mma_kernel<NB8, false, MODE, MINB, BATCH_CT>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    k<<<blocks, NTHR, smem, st>>>(a);
  }
  return check_launch("bnn_mma_kernel");
}

template <int MODE>
static int launch_mma_batch(const BnnArgs& a, cudaStream_t st) {
  switch ((a.batch + 7) / 8) {
    case 1: return launch_mma<1, MODE>(a, st);
    case 2: return launch_mma<2, MODE>(a, st);
    case 3: return a.batch == 20 ? launch_mma<3, MODE, 6, 20>(a, st)      // the reference's minibatch: masks fold
                                 : launch_mma<3, MODE>(a, st);
    default: return launch_mma<4, MODE>(a, st);
  }
}

static int launch_nll_grad(const BnnArgs& a, cudaStream_t st) {
  if (g_bnn_variant >= 10 && a.batch <= 32) {
    switch (g_bnn_variant) {              // accuracy modes of the tensor-pipe kernel (bnn_mma.cuh)
      case 11: return launch_mma_batch<MMA_ROUND_SPLIT>(a, st);
      case 12: return launch_mma_batch<MMA_RN_ACCUM>(a, st);
      case 13: return launch_mma_batch<MMA_ROUND_SPLIT | MMA_RN_ACCUM>(a, st);
      case 14: return launch_mma_batch<MMA_ROUND_SPLIT | MMA_RN_ACCUM | MMA_PACKED_SPLIT>(a, st);
      case 15: return launch_mma_batch<MMA_ROUND_SPLIT | MMA_RN_ACCUM | MMA_SEP_CROSS>(a, st);
      case 16: return launch_mma_batch<MMA_ROUND_SPLIT | MMA_RN_ACCUM | MMA_SEP_CROSS | MMA_PACKED_SPLIT>(a, st);
      default: return launch_mma_batch<0>(a, st);
    }
  }
  static const int nc_of_variant[10] = {5, 5, 8, 8, 5, 12, 8, 4, 4, 4};
  const int variant = g_bnn_variant < 10 ? g_bnn_variant : 0;
  const size_t per_chain = bnn_smem_floats(a.batch, a.L.n_in) * sizeof(float);
  if (per_chain * nc_of_variant[variant] > 227 * 1024) return launch_variant<1, 50, 1, 4, 1, 1>(a, st);
  switch (variant) {
    case 1: return launch_variant<1, 50, 5, 4, 1, 3>(a, st);     // 1 unit / thread, 15 chains / SM
    case 2: return launch_variant<2, 25, 8, 4, 2, 2>(a, st);     // 2 units / thread, 16 chains / SM
    case 3: return launch_variant<2, 25, 8, 2, 2, 2>(a, st);     //   ... 2 rows in flight
    case 4: return launch_variant<2, 25, 5, 4, 2, 3>(a, st);     //   ... 15 chains / SM in 3 CTAs
    case 5: return launch_variant<3, 17, 12, 4, 3, 1>(a, st);    // 3 units / thread, 12 chains / SM
    case 6: return launch_variant<3, 17, 8, 4, 3, 2>(a, st);     //   ... 16 chains / SM
    case 7: return launch_variant<2, 32, 4, 4, 2, 3>(a, st);     // one warp per chain, 12 chains / SM
    case 8: return launch_variant<2, 32, 4, 4, 2, 4>(a, st);     //   ... 16 chains / SM (<= 128 registers)
    case 9: return launch_variant<1, 50, 4, 4, 1, 4>(a, st);     // 1 unit / thread, 16 chains / SM
    default: return launch_variant<1, 50, 5, 4, 1, 2>(a, st);    // 1 unit / thread, 10 chains / SM
  }
}

// ---- two-stream pipeline of K5 (sgmcmc_set_bnn_pipeline) --------------------------------
// K4 of chunk j+1 (stream of the caller) runs WHILE K1 of chunk j (library-owned stream)
// streams that chunk's state: the SMs host CTAs of both kernels at once (both ask for the
// maximum shared-memory carveout, otherwise an SM would have to drain to switch), K1 fills the
// partial waves of K4 and vice versa, and a chunk's theta and gradient are still in L2 when
// K1 reads them.  The gradient goes through a ring of `ring` chunk-sized slots of grad_scratch,
// so its dirty lines are overwritten in L2 instead of being written back.
static int64_t g_pipe_chunk = 0;        // chains per chunk (0: pipeline off)
static int g_pipe_ring = 2;
void set_bnn_pipeline(int64_t chunk, int ring) { g_pipe_chunk = chunk; g_pipe_ring = ring; }

constexpr int PIPE_MAX_RING = 8;
struct PipeResources {
  int device = -1;
  cudaStream_t s_upd = nullptr;
  cudaEvent_t k4_done[PIPE_MAX_RING], slot_free[PIPE_MAX_RING], begin = nullptr, snap = nullptr;
};
static PipeResources g_pipe[16];

static int get_pipe(PipeResources** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= 16) return set_error(SGMCMC_E_CUDA, "bnn pipeline: cudaGetDevice failed");
  PipeResources& r = g_pipe[dev];
  if (r.device != dev) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);      // hi = greatest priority (numerically lowest)
    e = cudaStreamCreateWithPriority(&r.s_upd, cudaStreamNonBlocking, hi);
    for (int i = 0; i < PIPE_MAX_RING && e == cudaSuccess; ++i) {
      e = cudaEventCreateWithFlags(&r.k4_done[i], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r.slot_free[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r.begin, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r.snap, cudaEventDisableTiming);
    if (e != cudaSuccess) return set_error(SGMCMC_E_CUDA, "bnn pipeline: %s", cudaGetErrorString(e));
    r.device = dev;
  }
  *out = &r;
  return SGMCMC_OK;
}

#define SG_CUDA_RC(call)                                                                     \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) return set_error(SGMCMC_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

static int make_bnn_args(BnnArgs& a, const float* theta, const float* X, const float* y,
                         const int32_t* starts, float* cost, float* grad, float* mse, int64_t n_chains,
                         int n_in, int batch, float batch_size_cfg, int64_t n_examples) {
  SG_REQUIRE(n_chains >= 0, SGMCMC_E_INVALID, "n_chains must be >= 0");
  SG_REQUIRE(theta && X && y && cost, SGMCMC_E_INVALID, "bnn: theta, X, y and cost must not be NULL");
  SG_REQUIRE(n_in >= 1 && n_in <= 64, SGMCMC_E_UNSUPPORTED, "bnn: n_in must be in [1, 64] (got %d)", n_in);
  SG_REQUIRE(batch >= 1 && batch <= 256, SGMCMC_E_UNSUPPORTED, "bnn: batch must be in [1, 256] (got %d)", batch);
  SG_REQUIRE(batch_size_cfg > 0 && n_examples >= 1, SGMCMC_E_INVALID, "bnn: batch_size_cfg and n_examples must be > 0");
  a.theta = theta; a.X = X; a.y = y; a.starts = starts; a.cost = cost; a.grad = grad; a.mse = mse;
  a.n_chains = n_chains; a.batch = batch;
  a.L = make_layout(n_in);
  SG_REQUIRE(aligned_to(theta, 8) && (a.L.D % 2 == 0), SGMCMC_E_ALIGN, "bnn: theta must be 8-byte aligned");
  a.inv_bs = 1.0f / batch_size_cfg;
  a.inv_n = 1.0f / (float)n_examples;
  a.prior_den_inv = 1.0f / ((float)a.L.D + 3e-16f);
  return SGMCMC_OK;
}

// trace[k] <- theta, cost_trace[k] <- cost (device-to-device, stream ordered)
static int snapshot(float* trace, float* cost_trace, int64_t k, const float* theta, const float* cost,
                    int64_t n_chains, int64_t D, cudaStream_t st) {
  cudaError_t e = cudaSuccess;
  if (trace != nullptr)
    e = cudaMemcpyAsync(trace + k * n_chains * D, theta, sizeof(float) * n_chains * D, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess && cost_trace != nullptr)
    e = cudaMemcpyAsync(cost_trace + k * n_chains, cost, sizeof(float) * n_chains, cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return set_error(SGMCMC_E_CUDA, "snapshot copy: %s", cudaGetErrorString(e));
  return SGMCMC_OK;
}

}  // namespace sgmcmc

using namespace sgmcmc;

extern "C" int sgmcmc_bnn_nll_grad_f32(const float* theta, const float* X, const float* y,
                                       const int32_t* starts, float* cost, float* grad, float* mse,
                                       int64_t n_chains, int n_in, int batch, float batch_size_cfg,
                                       int64_t n_examples, void* stream) {
  BnnArgs a;
  if (int rc = make_bnn_args(a, theta, X, y, starts, cost, grad, mse, n_chains, n_in, batch,
                             batch_size_cfg, n_examples))
    return rc;
  if (n_chains == 0) return SGMCMC_OK;
  return launch_nll_grad(a, (cudaStream_t)stream);
}

// K5: `n_steps` steps of BNN-SGHMC for all chains, driven from C with no host
// synchronisation: per step K4 (cost + gradient at the old theta, minibatch
// starts[s, :]) then K1 (fused SGHMC update), plus a device-to-device snapshot of
// (theta, cost) every keep_every-th step.  The gradient goes through the caller's
// `grad_scratch` [C, D].  Why this is two kernels and not one: DESIGN.md "K5".
extern "C" int sgmcmc_bnn_sghmc_run_f32(float* theta, float* v, float* tau, float* g, float* v_hat,
                                        float* minv, const float* X, const float* y,
                                        const int32_t* starts, const float* z, float* trace,
                                        float* cost_trace, float* grad_scratch, float* cost_scratch,
                                        int64_t n_chains, int n_in, int batch,
                                        float batch_size_cfg, int64_t n_examples, int64_t n_steps,
                                        int64_t n_burn_in, int adapt_forever, int64_t keep_every,
                                        float epsilon, float mdecay, float scale_grad, uint64_t seed,
                                        uint64_t step0, uint64_t chain_offset, void* stream) {
  SG_REQUIRE(n_steps >= 0 && n_burn_in >= 0 && keep_every >= 1, SGMCMC_E_INVALID,
             "bnn_sghmc_run: n_steps, n_burn_in must be >= 0 and keep_every >= 1");
  SG_REQUIRE(cost_scratch, SGMCMC_E_INVALID, "bnn_sghmc_run: cost_scratch must not be NULL");
  SG_REQUIRE(v && tau && g && v_hat && minv, SGMCMC_E_INVALID, "bnn_sghmc_run: state arrays must not be NULL");
  BnnArgs a;
  if (int rc = make_bnn_args(a, theta, X, y, starts, cost_scratch, grad_scratch, nullptr, n_chains, n_in,
                             batch, batch_size_cfg, n_examples))
    return rc;
  const int64_t D = a.L.D, n = n_chains * D;
  SG_REQUIRE((chain_offset * (uint64_t)D) % 4 == 0, SGMCMC_E_INVALID, "chain_offset * D must be a multiple of 4");
  if (n_chains == 0) return SGMCMC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // Chains are processed in chunks: K4 then K1 on the same chunk back to back, with the
  // gradient of every chunk going through the SAME chunk-sized part of grad_scratch.  A
  // chunk's gradient (and its theta) is then still in the 126 MB L2 when K1 reads it, and the
  // next chunk overwrites the dirty gradient lines before they are written back, so the
  // gradient never costs HBM bandwidth (52 -> 40 B per element-step).
  const int64_t chunk = g_bnn_chunk > 0 ? g_bnn_chunk : n_chains;
  FusedStepArgs f;
  f.theta = theta; f.v = v; f.tau = tau; f.g = g; f.v_hat = v_hat; f.minv = minv;
  f.z = z;
  f.s = make_sghmc_scalars<float>(epsilon, mdecay, scale_grad);
  f.burn_in = 1; f.store_minv = 0;
  SG_REQUIRE(scale_grad > 0, SGMCMC_E_INVALID, "bnn_sghmc_run: scale_grad must be > 0");
  // one kernel per step (bnn_fused.cu) whenever the shape allows it; else K4 then K1
  const bool fused = bnn_fused_enabled() && g_bnn_chunk == 0 && bnn_fused_supported(a, f);
  SG_REQUIRE(fused || grad_scratch, SGMCMC_E_INVALID, "bnn_sghmc_run: grad_scratch must not be NULL");
  if (!fused && g_pipe_chunk > 0 && g_pipe_chunk < n_chains && z == nullptr) {
    // ---- pipelined: K4 chunks on `st`, K1 chunks on the library's update stream -----------
    PipeResources* P = nullptr;
    if (int rc = get_pipe(&P)) return rc;
    const int64_t pc = g_pipe_chunk;
    int ring = g_pipe_ring < 1 ? 1 : (g_pipe_ring > PIPE_MAX_RING ? PIPE_MAX_RING : g_pipe_ring);
    const int64_t n_chunks = (n_chains + pc - 1) / pc;
    if (ring > n_chains / pc) ring = (int)(n_chains / pc);      // the ring lives inside grad_scratch [C, D]
    SG_REQUIRE((pc * D) % 4 == 0, SGMCMC_E_INVALID, "bnn pipeline: chunk * D must be a multiple of 4");
    SG_CUDA_RC(cudaEventRecord(P->begin, st));
    SG_CUDA_RC(cudaStreamWaitEvent(P->s_upd, P->begin, 0));          // everything queued before this call
    int64_t j = 0;                                                   // global chunk counter
    for (int64_t s = 0; s < n_steps; ++s) {
      const int burn_in = adapt_forever || s < n_burn_in;
      const int store_minv = burn_in && (s == n_burn_in - 1 || (adapt_forever && s == n_steps - 1));
      for (int64_t c0 = 0; c0 < n_chains; c0 += pc, ++j) {
        const int64_t nc = n_chains - c0 < pc ? n_chains - c0 : pc;
        const int slot = (int)(j % ring);
        float* gslot = grad_scratch + (int64_t)slot * pc * D;
        // K1 of chunk j - ring has read this slot; it comes after K1 of this chunk's previous
        // step on the in-order update stream, so theta of the chunk is up to date as well
        if (j >= ring) SG_CUDA_RC(cudaStreamWaitEvent(st, P->slot_free[slot], 0));
        BnnArgs ac = a;
        ac.theta = theta + c0 * D;
        ac.cost = cost_scratch + c0;
        ac.grad = gslot;
        ac.n_chains = nc;
        ac.starts = starts != nullptr ? starts + s * n_chains + c0 : nullptr;
        if (int rc = launch_nll_grad(ac, st)) return rc;
        SG_CUDA_RC(cudaEventRecord(P->k4_done[slot], st));
        SG_CUDA_RC(cudaStreamWaitEvent(P->s_upd, P->k4_done[slot], 0));
        const int64_t o = c0 * D;
        if (int rc = sgmcmc_sghmc_step_f32(theta + o, v + o, tau + o, g + o, v_hat + o, minv + o, gslot, nullptr,
                                           nc * D, epsilon, mdecay, scale_grad, burn_in, store_minv, seed,
                                           step0 + (uint64_t)s, (chain_offset + (uint64_t)c0) * (uint64_t)D,
                                           (void*)P->s_upd))
          return rc;
        SG_CUDA_RC(cudaEventRecord(P->slot_free[slot], P->s_upd));
      }
      if ((s + 1) % keep_every == 0) {
        // the sample and the costs of this step, after its last K1 and before the next K4
        // overwrites the costs
        if (int rc = snapshot(trace, cost_trace, (s + 1) / keep_every - 1, theta, cost_scratch, n_chains, D,
                              P->s_upd))
          return rc;
        SG_CUDA_RC(cudaEventRecord(P->snap, P->s_upd));
        SG_CUDA_RC(cudaStreamWaitEvent(st, P->snap, 0));
      }
    }
    // the caller's stream continues after the last update
    SG_CUDA_RC(cudaEventRecord(P->snap, P->s_upd));
    SG_CUDA_RC(cudaStreamWaitEvent(st, P->snap, 0));
    return SGMCMC_OK;
  }
  for (int64_t s = 0; s < n_steps; ++s) {
    const int burn_in = adapt_forever || s < n_burn_in;
    const int store_minv = burn_in && (s == n_burn_in - 1 || (adapt_forever && s == n_steps - 1));
    if (fused) {
      BnnArgs ac = a;
      ac.starts = starts != nullptr ? starts + s * n_chains : nullptr;
      f.z = z != nullptr ? z + s * n : nullptr;
      f.na = NoiseArgs{seed, step0 + (uint64_t)s, chain_offset * (uint64_t)D / 4};
      f.burn_in = burn_in; f.store_minv = store_minv;
      if (int rc = launch_bnn_sghmc_fused(ac, f, st)) return rc;
      if ((s + 1) % keep_every == 0)
        if (int rc = snapshot(trace, cost_trace, (s + 1) / keep_every - 1, theta, cost_scratch, n_chains, D, st))
          return rc;
      continue;
    }
    for (int64_t c0 = 0; c0 < n_chains; c0 += chunk) {
      const int64_t nc = n_chains - c0 < chunk ? n_chains - c0 : chunk;
      BnnArgs ac = a;
      ac.theta = theta + c0 * D;
      ac.cost = cost_scratch + c0;
      ac.grad = grad_scratch;
      ac.n_chains = nc;
      ac.starts = starts != nullptr ? starts + s * n_chains + c0 : nullptr;
      if (int rc = launch_nll_grad(ac, st)) return rc;
      const int64_t o = c0 * D;
      if (int rc = sgmcmc_sghmc_step_f32(theta + o, v + o, tau + o, g + o, v_hat + o, minv + o, grad_scratch,
                                         z != nullptr ? z + s * n + o : nullptr, nc * D, epsilon, mdecay,
                                         scale_grad, burn_in, store_minv, seed, step0 + (uint64_t)s,
                                         (chain_offset + (uint64_t)c0) * (uint64_t)D, stream))
        return rc;
    }
    if ((s + 1) % keep_every == 0)
      if (int rc = snapshot(trace, cost_trace, (s + 1) / keep_every - 1, theta, cost_scratch, n_chains, D, st))
        return rc;
  }
  return SGMCMC_OK;
}

extern "C" int sgmcmc_bnn_predict_f32(const float* theta, const float* X, float* out, int64_t n_nets,
                                      int n_in, int64_t n_points, void* stream) {
  SG_REQUIRE(n_nets >= 0 && n_points >= 0, SGMCMC_E_INVALID, "negative size");
  if (n_nets == 0 || n_points == 0) return SGMCMC_OK;
  SG_REQUIRE(theta && X && out, SGMCMC_E_INVALID, "bnn_predict: NULL pointer");
  SG_REQUIRE(n_in >= 1 && n_in <= 64, SGMCMC_E_UNSUPPORTED, "bnn: n_in must be in [1, 64] (got %d)", n_in);
  const BnnLayout L = make_layout(n_in);
  using Shape = BnnShape<K10_COLS, K10_TPC, K10_NC>;
  const size_t smem = (size_t)K10_NC * bnn_smem_floats(PB, n_in) * sizeof(float);
  SG_REQUIRE(smem <= 227 * 1024, SGMCMC_E_UNSUPPORTED, "bnn_predict: n_in too large for shared memory");
  auto k = bnn_predict_kernel<K10_COLS, K10_TPC, K10_NC>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int64_t items = n_nets * ((n_points + PB - 1) / PB);
  k<<<(unsigned)((items + K10_NC - 1) / K10_NC), Shape::THREADS, smem, (cudaStream_t)stream>>>(
      theta, X, out, n_nets, n_points, L);
  return check_launch("bnn_predict_kernel");
}
