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
#include "mlp_umma.cuh"

namespace sgmcmc {

// wide layers on the tensor cores (csrc/mlp_umma.cu); sgmcmc_set_mlp_tuning(0) keeps every layer on
// the FFMA kernels below (the second implementation the tests compare against)
static int g_mlp_umma = 1;
void set_mlp_umma(int on) { g_mlp_umma = on; }

constexpr int MLP_MAX_W = 8;        // weight matrices: up to 7 hidden layers + the head
constexpr int MLP_THREADS = 128;
constexpr int MLP_KMAX = 512;       // rows of the staged operand per shared-memory chunk (40 KB at 20 rows)

// batch tile: minibatch rows padded to what the register tiles are compiled for
__host__ __device__ __forceinline__ int mlp_bt(int batch) { return batch <= 8 ? 8 : batch <= 16 ? 16 : batch <= 20 ? 20 : 32; }

struct MlpLayout {
  int n_w;                          // weight matrices = hidden layers + 1
  int width[MLP_MAX_W + 1];         // width[0] = n_in, width[1..n_w-1] hidden, width[n_w] = 1
  int64_t oW[MLP_MAX_W], ob[MLP_MAX_W], orho, D;
  // workspace (floats per chain).  Activations and their gradients are stored TRANSPOSED,
  // Ht_l[i * BT + b] = H_l[b][i] (BT = mlp_bt(batch), rows b >= batch are zero): the slice of
  // rows a GEMM chunk needs is then one contiguous block, staged into shared memory with
  // 128-bit copies and read back as 128-bit warp-broadcast loads (4 batch rows per load).
  int64_t oH[MLP_MAX_W], oZ[MLP_MAX_W];   // Ht_l, dZt_l for l = 1..n_w-1
  int64_t oSq, ws_floats;           // partial sums of squares (weight prior); total per chain
  int sq_slot[MLP_MAX_W];           // first partial-sum slot of layer l (one slot per column tile)
  int sq_n[MLP_MAX_W];              // partial sums layer l writes
  int n_sq;
  // tensor-core layers (mlp_umma.cu): umma[l - 1] != 0 when weight matrix l runs its forward and
  // backward-data GEMMs there; their activation operands live a second time in the workspace, split
  // into hi / lo planes in the MMA's canonical order (mlp_umma.cuh): Hc_l = H_l, Zc_l = dZ_l
  int umma[MLP_MAX_W];
  int64_t oHc[MLP_MAX_W], oZc[MLP_MAX_W], cplane[MLP_MAX_W];   // -1: not kept
};

struct MlpArgs {
  const float* theta;      // [n_theta_rows, D]
  const float* X;          // [n_rows, n_in]
  const float* y;          // [n_rows] (NULL for predict)
  const int32_t* starts;   // [C] first row of each chain's minibatch, or NULL (see item_rows)
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

// columns per thread of the column-owner kernels (forward, weight gradient) for a layer of
// n_out units, and the number of column tiles (CTAs per chain) that gives
__host__ __device__ __forceinline__ int mlp_cpt(int n_out) { return n_out >= 256 ? 4 : 1; }
__host__ __device__ __forceinline__ int mlp_col_tiles(int n_out) {
  return (n_out + mlp_cpt(n_out) * MLP_THREADS - 1) / (mlp_cpt(n_out) * MLP_THREADS);
}

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
  }
  for (int l = 0; l < L.n_w; ++l) {
    L.oW[l] = o; o += (int64_t)L.width[l] * L.width[l + 1];
    L.ob[l] = o; o += L.width[l + 1];
  }
  L.orho = o; o += 1;
  L.D = o;
  for (int l = L.n_w; l >= 1; --l) {                              // weight matrix l: width[l-1] -> width[l]
    const int wi = L.width[l - 1], wo = L.width[l];
    // tensor-core layers: widths that fill a 128-unit tile, rows of W_l that 64-bit loads can walk
    const bool ok = g_mlp_umma && l >= 2 && l < L.n_w && wi >= 128 && wo >= 128 && wi % 4 == 0 && wo % 4 == 0 &&
                    L.D % 2 == 0 && L.oW[l - 1] % 2 == 0;
    // bit 0: forward GEMM, bit 1: backward-data GEMM.  The backward one needs dZ_l split and in canonical
    // order, which the head and the tensor-core backward of layer l + 1 write (the FFMA backward does not)
    const bool bwd = ok && (l == L.n_w - 1 || (L.umma[l] & 2) != 0);
    L.umma[l - 1] = (ok ? 1 : 0) | (bwd ? 2 : 0);
  }
  for (int l = 0; l < L.n_w; ++l) {
    L.sq_slot[l] = slot;
    L.sq_n[l] = (L.umma[l] & 1) ? (L.width[l + 1] + 127) / 128 : mlp_col_tiles(L.width[l + 1]);
    slot += (L.width[l + 1] + MLP_THREADS - 1) / MLP_THREADS;     // >= the column tiles of either CPT / the 128-unit tiles
  }
  L.n_sq = slot;
  const int bt = mlp_bt(batch);
  L.oH[0] = L.oZ[0] = 0;
  L.oHc[0] = L.oZc[0] = -1;
  for (int l = 1; l < L.n_w; ++l) {
    L.oH[l] = w; w += (int64_t)bt * L.width[l];
    L.oZ[l] = w; w += (int64_t)bt * L.width[l];
    L.cplane[l] = (int64_t)((L.width[l] + MU_KPAD - 1) & ~(MU_KPAD - 1)) * MU_CN;
    L.oHc[l] = L.oZc[l] = -1;
    if (l + 1 < L.n_w && (L.umma[l] & 1)) { L.oHc[l] = w; w += 2 * L.cplane[l]; }    // operand of layer l+1's forward
    if (L.umma[l - 1] & 2) { L.oZc[l] = w; w += 2 * L.cplane[l]; }                  // operand of layer l's backward
  }
  L.oSq = w; w += (slot + 3) & ~3;
  L.ws_floats = (w + 3) & ~3;
  return SGMCMC_OK;
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

// Stage rows [k0, k0 + kc) of the left operand of layer l into shared memory as sH[ii * BT + b]:
// a contiguous 128-bit copy of Ht_{l-1} (l > 1), or the minibatch transposed (l = 1).
template <int BT>
__device__ __forceinline__ void stage_left(const MlpArgs& a, int l, const float* __restrict__ ws, int64_t row0,
                                           int rows, int k0, int kc, float* __restrict__ sH) {
  if (l > 1) {
    const float4* src = reinterpret_cast<const float4*>(ws + a.L.oH[l - 1] + (int64_t)k0 * BT);
    float4* dst = reinterpret_cast<float4*>(sH);
    for (int e = threadIdx.x; e < kc * (BT / 4); e += blockDim.x) dst[e] = src[e];
  } else {
    const int n_in = a.L.width[0];
    for (int e = threadIdx.x; e < kc * BT; e += blockDim.x) {
      const int b = e / kc, ii = e - b * kc;                     // ii fastest: coalesced global reads
      sH[ii * BT + b] = b < rows ? __ldg(a.X + (row0 + b) * n_in + k0 + ii) : 0.0f;
    }
  }
}

// CPT weights of one row, loaded with the widest access the alignment allows (VW floats per load)
template <int CPT, int VW>
__device__ __forceinline__ void load_w(const float* __restrict__ p, int n_valid, float (&w)[CPT]) {
  if constexpr (CPT == 4 && VW == 4) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
  } else if constexpr (CPT == 4 && VW == 2) {
    const float2 q0 = __ldg(reinterpret_cast<const float2*>(p)), q1 = __ldg(reinterpret_cast<const float2*>(p) + 1);
    w[0] = q0.x; w[1] = q0.y; w[2] = q1.x; w[3] = q1.y;
  } else {
#pragma unroll
    for (int c = 0; c < CPT; ++c) w[c] = c < n_valid ? __ldg(p + c) : 0.0f;
  }
}
// widest vector access for CPT columns starting at a multiple of CPT in rows of n_out floats
template <int CPT>
__device__ __forceinline__ int vector_width(const float* base, int n_out) {
  if (CPT != 4) return 1;
  if ((n_out & 3) == 0 && aligned_to_dev(base, 16)) return 4;
  if ((n_out & 1) == 0 && aligned_to_dev(base, 8)) return 2;      // (odd chains of a network with D % 4 == 2)
  return 1;
}

// ---- forward: H_l = tanh(H_{l-1} W_l + b_l), thread = CPT consecutive columns x all rows ----------
template <int BT, int CPT, int VW>
__device__ __forceinline__ void fwd_rows(const float* __restrict__ wrow, int n_out, int n_valid, int kc,
                                         const float* __restrict__ sH, float (&acc)[BT][CPT], float& sq) {
#pragma unroll 8
  for (int ii = 0; ii < kc; ++ii) {
    float w[CPT];
    load_w<CPT, VW>(wrow + (int64_t)ii * n_out, n_valid, w);
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

// The split, canonically ordered copy of a layer's output for a tensor-core GEMM (mlp_umma.cuh): units
// j0 .. j0 + CPT - 1 of this thread, all MU_CN rows (rows >= `rows` and units >= n_out are zeros).
// ACT: the values are pre-activations, apply bias + tanh (forward); else they are taken as they are.
template <int BT, int CPT>
__device__ __forceinline__ void store_canonical(float* __restrict__ dst, int64_t plane, const float (&acc)[BT][CPT],
                                                int j0, int n_out, int rows, bool act, const float* __restrict__ bias) {
  const int n_pad = (n_out + MU_KPAD - 1) & ~(MU_KPAD - 1);
  if (j0 >= n_pad) return;
  float bj[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) bj[c] = (act && j0 + c < n_out) ? __ldg(bias + j0 + c) : 0.0f;
  // a quad of units (j0 is a multiple of CPT) is one 16-byte piece per minibatch row: 128-bit stores
  float* q = dst + ((int64_t)(j0 >> 2) * MU_CN) * 4 + (j0 & 3);
#pragma unroll
  for (int n = 0; n < MU_CN; ++n) {
    float hi[CPT], lo[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      float v = 0.0f;
      if (n < BT) {
        v = acc[n < BT ? n : 0][c];
        if (act) v = fast_tanh(v + bj[c]);
        v = (n < rows && j0 + c < n_out) ? v : 0.0f;
      }
      hi[c] = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
      lo[c] = v - hi[c];
    }
    if constexpr (CPT == 4) {
      *reinterpret_cast<float4*>(q + 4 * n) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(q + plane + 4 * n) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    } else {
#pragma unroll
      for (int c = 0; c < CPT; ++c)
        if (j0 + c < n_pad) { q[4 * n + c] = hi[c]; q[plane + 4 * n + c] = lo[c]; }
    }
  }
}

template <int BT, int CPT>
__global__ void __launch_bounds__(MLP_THREADS) mlp_fwd_kernel(MlpArgs a, int l) {
  extern __shared__ __align__(16) float sH[];              // [min(n_in, MLP_KMAX)][BT]
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
  const int n_valid = n_out - j0;                          // columns of this thread that exist (<= 0: none)
  const int vw = vector_width<CPT>(W, n_out);
  float acc[BT][CPT];
#pragma unroll
  for (int b = 0; b < BT; ++b)
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[b][c] = 0.0f;
  float sq = 0.0f;
  for (int k0 = 0; k0 < n_in; k0 += MLP_KMAX) {
    const int kc = min(MLP_KMAX, n_in - k0);
    if (k0 > 0) __syncthreads();
    stage_left<BT>(a, l, ws, row0, rows, k0, kc, sH);
    __syncthreads();
    if (n_valid > 0) {
      const float* __restrict__ wrow = W + (int64_t)k0 * n_out + j0;
      if (vw == 4) fwd_rows<BT, CPT, 4>(wrow, n_out, n_valid, kc, sH, acc, sq);
      else if (vw == 2) fwd_rows<BT, CPT, 2>(wrow, n_out, n_valid, kc, sH, acc, sq);
      else fwd_rows<BT, CPT, 1>(wrow, n_out, n_valid, kc, sH, acc, sq);
    }
  }
  if (n_valid > 0) {
    float* __restrict__ Hout = ws + a.L.oH[l];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      if (c < n_valid) {
        const float bj = __ldg(bias + j0 + c);
        sq = fmaf(bj, bj, sq);
        float4* out4 = reinterpret_cast<float4*>(Hout + (int64_t)(j0 + c) * BT);
#pragma unroll
        for (int b4 = 0; b4 < BT / 4; ++b4) {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = 4 * b4 + e < rows ? fast_tanh(acc[4 * b4 + e][c] + bj) : 0.0f;
          out4[b4] = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
  }
  if (a.L.oHc[l] >= 0) store_canonical<BT, CPT>(ws + a.L.oHc[l], a.L.cplane[l], acc, j0, n_out, rows, true, bias);
  // sum of squares of this tile's weights and biases, for the weight prior (:131-141)
  const float tot = block_sum(sq, red);
  if (threadIdx.x == 0) ws[a.L.oSq + a.L.sq_slot[l - 1] + ct] = tot;
}

// ---- head: f = H_L W_{L+1} + b_{L+1}; loss (bayesian_neural_network.py:368-388); gradient of the
// head parameters and rho; dZ_L = (df W_{L+1}^T) * (1 - H_L^2).  One CTA per work item. -------------
template <int BT, bool WANT_GRAD>
__global__ void __launch_bounds__(MLP_THREADS) mlp_head_kernel(MlpArgs a) {
  __shared__ float sF[MLP_THREADS / 32][BT], sDf[BT], red[MLP_THREADS / 32];
  const int l = a.L.n_w;                       // the head is weight matrix n_w (1-based)
  const int hL = a.L.width[l - 1];
  const int64_t chain = blockIdx.x;
  const int64_t trow = chain / a.theta_div;
  const float* __restrict__ th = a.theta + trow * a.L.D;
  const float* __restrict__ W = th + a.L.oW[l - 1];
  float* __restrict__ ws = a.ws + chain * a.L.ws_floats;
  const float* __restrict__ Ht = ws + a.L.oH[l - 1];
  int64_t row0;
  int rows;
  item_rows(a, chain, row0, rows);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float b4 = th[a.L.ob[l - 1]];
  const float rho = th[a.L.orho];
  // f[b] = sum_i Ht[i][b] W[i]: a thread walks its units i, all rows at once
  float f[BT];
#pragma unroll
  for (int b = 0; b < BT; ++b) f[b] = 0.0f;
  float sq = 0.0f;
  for (int i = tid; i < hL; i += MLP_THREADS) {
    const float wi = __ldg(W + i);
    sq = fmaf(wi, wi, sq);
    const float4* h4 = reinterpret_cast<const float4*>(Ht + (int64_t)i * BT);
#pragma unroll
    for (int b4 = 0; b4 < BT / 4; ++b4) {
      const float4 h = h4[b4];
      f[4 * b4 + 0] = fmaf(h.x, wi, f[4 * b4 + 0]); f[4 * b4 + 1] = fmaf(h.y, wi, f[4 * b4 + 1]);
      f[4 * b4 + 2] = fmaf(h.z, wi, f[4 * b4 + 2]); f[4 * b4 + 3] = fmaf(h.w, wi, f[4 * b4 + 3]);
    }
  }
#pragma unroll
  for (int b = 0; b < BT; ++b) {
    float v = f[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sF[w][b] = v;
  }
  const float sq_head = block_sum(sq, red);                           // (its barriers also publish sF)
  const float e_rho = expf(rho);
  const float fvi = 1.0f / (e_rho + 1e-16f);                          // :368
  float sse = 0.0f;
  if (tid < 32) {
    float diff = 0.0f;
    if (tid < rows) {
      float fb = b4;
#pragma unroll
      for (int ww = 0; ww < MLP_THREADS / 32; ++ww) fb += sF[ww][tid];
      if (a.fout != nullptr) {                 // predict: (mean, log variance) per row (:535-557)
        float* o = a.fout + (trow * a.n_rows + row0 + tid) * 2;
        o[0] = fb;
        o[1] = rho;
      }
      if (a.y != nullptr) {
        diff = __ldg(a.y + row0 + tid) - fb;
        sDf[tid] = -(diff * fvi) * a.inv_bs;                          // d cost / d f_i
      }
    } else if (tid < BT) {
      sDf[tid] = 0.0f;
    }
    sse = diff * diff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sse += __shfl_xor_sync(0xffffffffu, sse, o);
  }
  if (a.y == nullptr) return;
  __syncthreads();
  const float pscale = a.prior_den_inv * a.inv_n;
  if (tid == 0) {
    float sq_t = sq_head + b4 * b4 + rho * rho;
    for (int q = 1; q < l; ++q)                                       // one partial per column tile of layer q
      for (int ct = 0; ct < a.L.sq_n[q - 1]; ++ct) sq_t += ws[a.L.oSq + a.L.sq_slot[q - 1] + ct];
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
  float* __restrict__ dZt = ws + a.L.oZ[l - 1];
  float df[BT];
#pragma unroll
  for (int b = 0; b < BT; ++b) df[b] = sDf[b];
  const int64_t oZc = a.L.oZc[l - 1], zplane = a.L.cplane[l - 1];    // tensor-core backward of layer L: split copy of dZ_L
  for (int i = tid; i < hL; i += MLP_THREADS) {
    const float wi = __ldg(W + i);
    const float4* h4 = reinterpret_cast<const float4*>(Ht + (int64_t)i * BT);
    float4* z4 = reinterpret_cast<float4*>(dZt + (int64_t)i * BT);
    float* zc = ws + oZc + ((int64_t)(i >> 2) * MU_CN) * 4 + (i & 3);
    float dw = 0.0f;
#pragma unroll
    for (int b4 = 0; b4 < BT / 4; ++b4) {
      const float4 h = h4[b4];
      const float hv[4] = {h.x, h.y, h.z, h.w};
      float zv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        dw = fmaf(hv[e], df[4 * b4 + e], dw);
        zv[e] = (df[4 * b4 + e] * wi) * fmaf(-hv[e], hv[e], 1.0f);    // (rows >= batch: df = 0)
        if (oZc >= 0) {
          const float hi = __uint_as_float((__float_as_uint(zv[e]) + 0x1000u) & 0xffffe000u);
          zc[4 * (4 * b4 + e)] = hi;
          zc[zplane + 4 * (4 * b4 + e)] = zv[e] - hi;
        }
      }
      z4[b4] = make_float4(zv[0], zv[1], zv[2], zv[3]);
    }
    if (oZc >= 0)
      for (int n = BT; n < MU_CN; ++n) zc[4 * n] = zc[zplane + 4 * n] = 0.0f;
    g[i] = fmaf(wi, pscale, dw);
  }
  if (oZc >= 0)                                                       // units hL .. round_up(hL, MU_KPAD): zeros
    for (int e = tid; e < (((hL + MU_KPAD - 1) & ~(MU_KPAD - 1)) - hL) * MU_CN; e += MLP_THREADS) {
      const int i = hL + e / MU_CN, n = e % MU_CN;
      float* zc = ws + oZc + ((int64_t)(i >> 2) * MU_CN + n) * 4 + (i & 3);
      zc[0] = zc[zplane] = 0.0f;
    }
}

// ---- weight gradient: dW_l[i][j] = sum_b H_{l-1}[b][i] dZ_l[b][j] + pscale W_l[i][j] (and db_l);
// thread = CPT consecutive columns (dZ of those columns in registers), the rows i of its slice ------
template <int BT, int CPT, int VW>
__device__ __forceinline__ void wgrad_rows(const float* __restrict__ W, float* __restrict__ gW, int n_out, int n_valid,
                                           int kc, const float* __restrict__ sH, const float (&dz)[BT][CPT],
                                           float pscale) {
#pragma unroll 8
  for (int ii = 0; ii < kc; ++ii) {
    const int64_t o = (int64_t)ii * n_out;
    float w[CPT], gsum[CPT];
    load_w<CPT, VW>(W + o, n_valid, w);
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
    if constexpr (CPT == 4 && VW == 4) {
      *reinterpret_cast<float4*>(gW + o) = make_float4(fmaf(w[0], pscale, gsum[0]), fmaf(w[1], pscale, gsum[1]),
                                                       fmaf(w[2], pscale, gsum[2]), fmaf(w[3], pscale, gsum[3]));
    } else if constexpr (CPT == 4 && VW == 2) {
      float2* g2 = reinterpret_cast<float2*>(gW + o);
      g2[0] = make_float2(fmaf(w[0], pscale, gsum[0]), fmaf(w[1], pscale, gsum[1]));
      g2[1] = make_float2(fmaf(w[2], pscale, gsum[2]), fmaf(w[3], pscale, gsum[3]));
    } else {
#pragma unroll
      for (int c = 0; c < CPT; ++c)
        if (c < n_valid) gW[o + c] = fmaf(w[c], pscale, gsum[c]);
    }
  }
}

template <int BT, int CPT>
__global__ void __launch_bounds__(MLP_THREADS, CPT == 4 ? 4 : 1) mlp_wgrad_kernel(MlpArgs a, int l, int n_is) {
  extern __shared__ __align__(16) float sH[];              // [min(rows of the slice, MLP_KMAX)][BT]
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
  const int n_valid = n_out - j0;
  const float pscale = a.prior_den_inv * a.inv_n;
  const int vw = min(vector_width<CPT>(W, n_out), vector_width<CPT>(gW, n_out));
  float dz[BT][CPT];
  const float* __restrict__ dZt = ws + a.L.oZ[l];
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    if (c < n_valid) {
      const float4* z4 = reinterpret_cast<const float4*>(dZt + (int64_t)(j0 + c) * BT);
#pragma unroll
      for (int b4 = 0; b4 < BT / 4; ++b4) {
        const float4 z = z4[b4];
        dz[4 * b4 + 0][c] = z.x; dz[4 * b4 + 1][c] = z.y; dz[4 * b4 + 2][c] = z.z; dz[4 * b4 + 3][c] = z.w;
      }
    } else {
#pragma unroll
      for (int b = 0; b < BT; ++b) dz[b][c] = 0.0f;
    }
  }
  if (is == 0) {
#pragma unroll
    for (int c = 0; c < CPT; ++c)
      if (c < n_valid) {
        float db = 0.0f;
#pragma unroll
        for (int b = 0; b < BT; ++b) db += dz[b][c];
        a.grad[chain * a.L.D + a.L.ob[l - 1] + j0 + c] = fmaf(__ldg(th + a.L.ob[l - 1] + j0 + c), pscale, db);
      }
  }
  const int ilen = (n_in + n_is - 1) / n_is;
  const int i_begin = is * ilen, i_end = min(n_in, i_begin + ilen);
  for (int k0 = i_begin; k0 < i_end; k0 += MLP_KMAX) {
    const int kc = min(MLP_KMAX, i_end - k0);
    if (k0 > i_begin) __syncthreads();
    stage_left<BT>(a, l, ws, row0, rows, k0, kc, sH);
    __syncthreads();
    if (n_valid > 0) {
      const int64_t o = (int64_t)k0 * n_out + j0;
      if (vw == 4) wgrad_rows<BT, CPT, 4>(W + o, gW + o, n_out, n_valid, kc, sH, dz, pscale);
      else if (vw == 2) wgrad_rows<BT, CPT, 2>(W + o, gW + o, n_out, n_valid, kc, sH, dz, pscale);
      else wgrad_rows<BT, CPT, 1>(W + o, gW + o, n_out, n_valid, kc, sH, dz, pscale);
    }
  }
}

// ---- backward data: dZ_{l-1}[b][i] = (sum_j dZ_l[b][j] W_l[i][j]) * (1 - H_{l-1}[b][i]^2);
// thread = RI rows i (interleaved, so a warp's rows are consecutive) x all batch rows -----------------
template <int BT, int RI>
__global__ void __launch_bounds__(MLP_THREADS) mlp_bwd_data_kernel(MlpArgs a, int l) {
  extern __shared__ __align__(16) float sZ[];              // [min(n_out, MLP_KMAX)][BT]
  const int n_in = a.L.width[l - 1], n_out = a.L.width[l];
  const int n_rt = (n_in + RI * MLP_THREADS - 1) / (RI * MLP_THREADS);
  const int64_t chain = blockIdx.x / n_rt;
  const int rt = (int)(blockIdx.x - chain * n_rt);
  const float* __restrict__ th = a.theta + (chain / a.theta_div) * a.L.D;
  const float* __restrict__ W = th + a.L.oW[l - 1];
  float* __restrict__ ws = a.ws + chain * a.L.ws_floats;
  int irow[RI];
#pragma unroll
  for (int r = 0; r < RI; ++r) irow[r] = (rt * RI + r) * MLP_THREADS + (int)threadIdx.x;
  // rows of W_l start at multiples of n_out floats: vector loads along a row need n_out % 4 == 0; a
  // base that is only 8-byte aligned (odd chains when D % 4 == 2) takes two 64-bit loads
  const bool vec = (n_out & 3) == 0 && aligned_to_dev(W, 8);
  const bool a16 = aligned_to_dev(W, 16);
  float acc[RI][BT];
#pragma unroll
  for (int r = 0; r < RI; ++r)
#pragma unroll
    for (int b = 0; b < BT; ++b) acc[r][b] = 0.0f;
  for (int k0 = 0; k0 < n_out; k0 += MLP_KMAX) {
    const int kc = min(MLP_KMAX, n_out - k0);
    if (k0 > 0) __syncthreads();
    {
      const float4* src = reinterpret_cast<const float4*>(ws + a.L.oZ[l] + (int64_t)k0 * BT);
      float4* dst = reinterpret_cast<float4*>(sZ);
      for (int e = threadIdx.x; e < kc * (BT / 4); e += blockDim.x) dst[e] = src[e];
    }
    __syncthreads();
    if (vec) {                                             // (kc is a multiple of 4: n_out and MLP_KMAX are)
#pragma unroll 2
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
#pragma unroll 4
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
  const float* __restrict__ Ht = ws + a.L.oH[l - 1];
  float* __restrict__ dZp = ws + a.L.oZ[l - 1];
  // (a layer whose backward runs on the tensor cores gets its dZ from mlp_umma.cu or the head, never from here:
  // see make_mlp_layout)
#pragma unroll
  for (int r = 0; r < RI; ++r)
    if (irow[r] < n_in) {
      const float4* h4 = reinterpret_cast<const float4*>(Ht + (int64_t)irow[r] * BT);
      float4* z4 = reinterpret_cast<float4*>(dZp + (int64_t)irow[r] * BT);
#pragma unroll
      for (int b4 = 0; b4 < BT / 4; ++b4) {
        const float4 h = h4[b4];                          // (rows >= batch: acc = 0, H = 0)
        z4[b4] = make_float4(acc[r][4 * b4 + 0] * fmaf(-h.x, h.x, 1.0f), acc[r][4 * b4 + 1] * fmaf(-h.y, h.y, 1.0f),
                             acc[r][4 * b4 + 2] * fmaf(-h.z, h.z, 1.0f), acc[r][4 * b4 + 3] * fmaf(-h.w, h.w, 1.0f));
      }
    }
}

// ---- host side ------------------------------------------------------------------------------------------
template <typename K>
static void allow_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <int BT>
static int launch_mlp_bt(const MlpArgs& a, bool want_grad, cudaStream_t st) {
  const MlpLayout& L = a.L;
  const int n_hidden = L.n_w - 1;
  const int64_t C = a.n_chains;
  auto smem_rows = [](int n) { return (size_t)(n < MLP_KMAX ? n : MLP_KMAX) * BT * sizeof(float); };
  auto umma_args = [&](int l, bool fwd) {        // weight matrix l on the tensor cores (mlp_umma.cu)
    MuArgs m;
    m.theta = a.theta; m.D = L.D; m.theta_div = a.theta_div;
    m.oW = L.oW[l - 1]; m.ob = L.ob[l - 1]; m.ldw = L.width[l];
    m.ws = a.ws; m.ws_floats = L.ws_floats;
    m.starts = a.starts; m.n_rows = a.n_rows; m.batch = a.batch; m.bt = BT;
    if (fwd) {                                   // H_l = tanh(H_{l-1} W_l + b_l)
      m.M = L.width[l]; m.K = L.width[l - 1];
      m.oBc = L.oHc[l - 1]; m.b_plane = L.cplane[l - 1];
      m.oHin = 0; m.oOut = L.oH[l];
      m.oOutc = L.oHc[l]; m.out_plane = L.cplane[l];
      m.oSq = L.oSq + L.sq_slot[l - 1];
    } else {                                     // dZ_{l-1} = (dZ_l W_l^T) * (1 - H_{l-1}^2)
      m.M = L.width[l - 1]; m.K = L.width[l];
      m.oBc = L.oZc[l]; m.b_plane = L.cplane[l];
      m.oHin = L.oH[l - 1]; m.oOut = L.oZ[l - 1];
      m.oOutc = L.oZc[l - 1]; m.out_plane = L.cplane[l - 1];
      m.oSq = 0;
    }
    m.Mpad = (m.M + MU_KPAD - 1) & ~(MU_KPAD - 1);
    return m;
  };
  for (int l = 1; l <= n_hidden; ++l) {
    if (L.umma[l - 1] & 1) {
      if (int rc = launch_mlp_gemm_umma(umma_args(l, true), true, C, st)) return rc;
      continue;
    }
    const int n_ct = mlp_col_tiles(L.width[l]);
    const size_t smem = smem_rows(L.width[l - 1]);
    if (mlp_cpt(L.width[l]) == 4) {
      allow_smem(mlp_fwd_kernel<BT, 4>, smem);
      mlp_fwd_kernel<BT, 4><<<(unsigned)(C * n_ct), MLP_THREADS, smem, st>>>(a, l);
    } else {
      allow_smem(mlp_fwd_kernel<BT, 1>, smem);
      mlp_fwd_kernel<BT, 1><<<(unsigned)(C * n_ct), MLP_THREADS, smem, st>>>(a, l);
    }
    if (int rc = check_launch("mlp_fwd_kernel")) return rc;
  }
  if (want_grad) mlp_head_kernel<BT, true><<<(unsigned)C, MLP_THREADS, 0, st>>>(a);
  else mlp_head_kernel<BT, false><<<(unsigned)C, MLP_THREADS, 0, st>>>(a);
  if (int rc = check_launch("mlp_head_kernel")) return rc;
  if (!want_grad) return SGMCMC_OK;
  for (int l = n_hidden; l >= 1; --l) {
    const int n_in = L.width[l - 1], n_out = L.width[l];
    // split the rows of W_l over CTAs until the grid fills the machine (few chains, wide layers)
    const bool quad = mlp_cpt(n_out) == 4;
    const int n_ct = mlp_col_tiles(n_out);
    int n_is = 1;
    while (C * n_ct * n_is < 592 && n_is * 2 * 64 <= n_in) n_is *= 2;
    const size_t smem = smem_rows((n_in + n_is - 1) / n_is);
    if (quad) {
      allow_smem(mlp_wgrad_kernel<BT, 4>, smem);
      mlp_wgrad_kernel<BT, 4><<<(unsigned)(C * n_ct * n_is), MLP_THREADS, smem, st>>>(a, l, n_is);
    } else {
      allow_smem(mlp_wgrad_kernel<BT, 1>, smem);
      mlp_wgrad_kernel<BT, 1><<<(unsigned)(C * n_ct * n_is), MLP_THREADS, smem, st>>>(a, l, n_is);
    }
    if (int rc = check_launch("mlp_wgrad_kernel")) return rc;
    if (l > 1 && (L.umma[l - 1] & 2)) {
      if (int rc = launch_mlp_gemm_umma(umma_args(l, false), false, C, st)) return rc;
    } else if (l > 1) {
      const size_t smem_z = smem_rows(n_out);
      const int n_rt4 = (n_in + 4 * MLP_THREADS - 1) / (4 * MLP_THREADS);
      if (BT <= 20 && n_in >= 4 * MLP_THREADS && C * n_rt4 >= 296) {
        allow_smem(mlp_bwd_data_kernel<BT, 4>, smem_z);
        mlp_bwd_data_kernel<BT, 4><<<(unsigned)(C * n_rt4), MLP_THREADS, smem_z, st>>>(a, l);
      } else {
        const int n_rt = (n_in + MLP_THREADS - 1) / MLP_THREADS;
        allow_smem(mlp_bwd_data_kernel<BT, 1>, smem_z);
        mlp_bwd_data_kernel<BT, 1><<<(unsigned)(C * n_rt), MLP_THREADS, smem_z, st>>>(a, l);
      }
      if (int rc = check_launch("mlp_bwd_data_kernel")) return rc;
    }
  }
  return SGMCMC_OK;
}

static bool uses_umma(const MlpLayout& L) {
  for (int l = 0; l < L.n_w; ++l)
    if (L.umma[l]) return true;
  return false;
}

static int launch_mlp(const MlpArgs& a, bool want_grad, cudaStream_t st) {
  switch (mlp_bt(a.batch)) {
    case 8: return launch_mlp_bt<8>(a, want_grad, st);
    case 16: return launch_mlp_bt<16>(a, want_grad, st);
    case 20: return launch_mlp_bt<20>(a, want_grad, st);
    default: return launch_mlp_bt<32>(a, want_grad, st);
  }
}

}  // namespace sgmcmc

using namespace sgmcmc;

extern "C" int sgmcmc_set_mlp_tuning(int tensor_core_layers) {
  set_mlp_umma(tensor_core_layers != 0);
  return SGMCMC_OK;
}

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
  SG_REQUIRE(!uses_umma(a.L) || aligned_to(theta, 8), SGMCMC_E_ALIGN, "mlp: theta must be 8-byte aligned");
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
  SG_REQUIRE(!uses_umma(a.L) || aligned_to(theta, 8), SGMCMC_E_ALIGN, "mlp: theta must be 8-byte aligned");
  a.theta = theta; a.X = X; a.y = nullptr; a.starts = nullptr; a.ws = (float*)workspace;
  a.cost = nullptr; a.grad = nullptr; a.mse = nullptr; a.fout = out;
  a.n_chains = n_nets * tiles; a.n_rows = n_points; a.theta_div = (int)tiles; a.batch = batch;
  a.inv_bs = 1.0f; a.inv_n = 1.0f; a.prior_den_inv = 1.0f;
  return launch_mlp(a, false, (cudaStream_t)stream);
}
