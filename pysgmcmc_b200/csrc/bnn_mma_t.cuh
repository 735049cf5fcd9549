// K4 on the tensor pipe, TRANSPOSED formulation (launch variants 14-17).
// Cost + gradient of one BOHAMIANN chain per CTA of 4 warps; replaces
// pysgmcmc/models/bayesian_neural_network.py:28-69,77-141,337-388 + tf.gradients, like bnn_mma.cuh.
//
// What round 1's tensor-pipe kernel (bnn_mma.cuh) paid for, per its ncu capture: the hi/lo split of
// every operand word at every use (26 % of the executed instructions -- each weight split by both
// warps, and again in the backward pass), 2-way bank conflicts of the stride-50 weight gathers, and
// (accuracy mode 13) one FADD per accumulator element and k-step.  Here:
//  * the weights are the M operand (out units 50 -> 64 = 4 m-tiles, one per warp), the minibatch the
//    N operand (20 -> 24 = 3 n-tiles): a warp only ever loads ITS OWN rows of a weight matrix;
//  * operands can be split ONCE, when they are produced (PRE_W: theta while it is staged; PRE_A: an
//    activation in the epilogue of the GEMM that made it) and then live in shared memory as a hi
//    and a lo plane: a fragment word is a load and no ALU instruction;
//  * the k-slot and row permutations an MMA is free to use (any bijection of the 8 k slots applied to
//    both operands; any assignment of accumulator rows to units) are chosen so that the registers of
//    a fragment that must be consecutive are ADJACENT IN MEMORY: (a0, a1) and (a2, a3) of the
//    forward / weight-gradient A fragments and (b0, b1) of every B fragment are one LDS.64 each, with
//    no register moves, and row strides 52 (weights) / 56 (activations) keep all of them -- and the
//    scalar gathers of the backward A fragment -- free of bank conflicts;
//  * accumulation across k-steps is done by the FP32 pipe with packed add.rn.f32x2 (two
//    accumulator elements per instruction): the tensor core truncates when it adds into its C
//    operand, which is what made the chained variants drift (tools/bnn_trajectory_drift.py).
// Arithmetic per product as in bnn_mma.cuh mode 13: x = hi + lo with hi = tf32_rn(x),
// a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi inside one k-step (zero C operand), k-steps added RN.
//
// Shared memory per chain (fp32 planes; the lo planes only with PRE_W / PRE_A):
//   W2, W3   [52][52]  [W; b] rows 0..50 (in unit / bias), row 51 and columns 50, 51 zero
//   H1, H2   [RB][56]  batch-major, unit-minor; unit 50 is the constant 1 that multiplies the bias row
//   Z        [RB][56]  dZ of the current layer, units stored at pos(u) = 8(u/8) + 2(u%4) + (u%8)/4 so
//                      that units u, u+4 are adjacent (the standard k slots t, t+4 of its consumer)
//   fp32: head and tail of theta (W1, b1 | W4, b4, rho), the minibatch, cross-warp scratch.
#pragma once

#include "bnn_mma.cuh"

namespace sgmcmc {

constexpr int TW = 52;          // row stride and rows of a weight plane
constexpr int TH = 56;          // row stride of an activation plane (units 0..49, the 1, 5 zeros)
constexpr int T_THREADS = 128;

__host__ __device__ constexpr int bnn_t_rows(int batch) { return batch <= 20 ? (batch <= 8 ? 8 : batch <= 16 ? 16 : 20) : batch <= 24 ? 24 : 32; }

template <bool PRE_W, bool PRE_A>
__host__ __device__ constexpr int bnn_t_smem_bytes(int rb, int nb8, int n_in) {
  // W2, W3 planes | H1, H2, Z planes | over-read pad | fp32 part
  return 4 * ((PRE_W ? 4 : 2) * TW * TW + (PRE_A ? 6 : 3) * rb * TH + 12 * TH +
              ((n_in + 1) * HID + 3) / 4 * 4 + 56 + 8 * nb8 * n_in + 8 * nb8 + 4 * 32 + 16);
}

__device__ __forceinline__ void split_rn(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = x - hi;
}

// acc(2) += part(2), round to nearest, one packed instruction
__device__ __forceinline__ void add2(float& a0, float& a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tadd.rn.f32x2 ra, ra, rb;\n\t"
      "mov.b64 {%0, %1}, ra;\n\t}"
      : "+f"(a0), "+f"(a1)
      : "f"(b0), "f"(b1));
}

// One k-step of a [16 x 8*NT] tile: part = a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, then acc (+)= part.
template <int NT, bool FIRST>
__device__ __forceinline__ void kstep(float (&acc)[NT][4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                      const uint32_t (&bh)[NT][2], const uint32_t (&bl)[NT][2]) {
  if constexpr (FIRST) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) mma_tf32_zero(acc[nt], al, bh[nt]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[nt], ah, bl[nt]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[nt], ah, bh[nt]);
  } else {
    float part[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) mma_tf32_zero(part[nt], al, bh[nt]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) mma_tf32(part[nt], ah, bl[nt]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) mma_tf32(part[nt], ah, bh[nt]);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      add2(acc[nt][0], acc[nt][1], part[nt][0], part[nt][1]);
      add2(acc[nt][2], acc[nt][3], part[nt][2], part[nt][3]);
    }
  }
}

// two adjacent operand words (one LDS.64 per plane) -> consecutive fragment registers (hi, lo)
template <bool PRE>
__device__ __forceinline__ void ld2(const float* __restrict__ p, int lo_off, uint32_t& h0, uint32_t& h1,
                                    uint32_t& l0, uint32_t& l1) {
  const float2 v = *reinterpret_cast<const float2*>(p);
  if constexpr (PRE) {
    const float2 u = *reinterpret_cast<const float2*>(p + lo_off);
    h0 = __float_as_uint(v.x); h1 = __float_as_uint(v.y);
    l0 = __float_as_uint(u.x); l1 = __float_as_uint(u.y);
  } else {
    float a, b, c, d;
    split_rn(v.x, a, b);
    split_rn(v.y, c, d);
    h0 = __float_as_uint(a); l0 = __float_as_uint(b);
    h1 = __float_as_uint(c); l1 = __float_as_uint(d);
  }
}
template <bool PRE>
__device__ __forceinline__ void ld1(const float* __restrict__ p, int lo_off, uint32_t& h, uint32_t& l) {
  const float v = *p;
  if constexpr (PRE) {
    h = __float_as_uint(v);
    l = __float_as_uint(p[lo_off]);
  } else {
    float a, b;
    split_rn(v, a, b);
    h = __float_as_uint(a); l = __float_as_uint(b);
  }
}

// position of unit u inside a row of the Z plane
__device__ __forceinline__ constexpr int zpos(int u) { return (u & ~7) + 2 * (u & 3) + ((u >> 2) & 1); }

// ---- forward GEMM:  acc[out unit][batch] = sum_k Wb[k][unit] * Hb[batch][k],  k = in unit (50: bias x 1)
// k slots: t <-> 8ks + 2t, t+4 <-> 8ks + 2t + 1;  rows: g <-> unit 16w + 2g, g+8 <-> 16w + 2g + 1;
// columns: batch 8nt + 2t (+1) as usual.  k = 52..55 (ks = 6, t >= 2) do not exist: constant zeros.
template <int NB8, bool PRE_W, bool PRE_A>
__device__ __forceinline__ void gemm_fwd(const float* __restrict__ W, const float* __restrict__ Hb, int a_lo, int b_lo,
                                         float (&acc)[NB8][4], int w, int g, int t) {
  const float* a_base = W + 2 * t * TW + 16 * w + 2 * g;
  const float* b_base = Hb + g * TH + 2 * t;
#pragma unroll
  for (int ks = 0; ks < 7; ++ks) {
    uint32_t ah[4], al[4], bh[NB8][2], bl[NB8][2];
    ld2<PRE_W>(a_base + 8 * ks * TW, a_lo, ah[0], ah[1], al[0], al[1]);
    ld2<PRE_W>(a_base + (8 * ks + 1) * TW, a_lo, ah[2], ah[3], al[2], al[3]);
#pragma unroll
    for (int nt = 0; nt < NB8; ++nt) ld2<PRE_A>(b_base + 8 * nt * TH + 8 * ks, b_lo, bh[nt][0], bh[nt][1], bl[nt][0], bl[nt][1]);
    if (ks == 6) {             // rows 52..55 of W / units 52..55 of H: t >= 2
      const bool dead = t >= 2;
#pragma unroll
      for (int e = 0; e < 4; ++e) { ah[e] = dead ? 0u : ah[e]; al[e] = dead ? 0u : al[e]; }
#pragma unroll
      for (int nt = 0; nt < NB8; ++nt) {
        bh[nt][0] = dead ? 0u : bh[nt][0]; bh[nt][1] = dead ? 0u : bh[nt][1];
        bl[nt][0] = dead ? 0u : bl[nt][0]; bl[nt][1] = dead ? 0u : bl[nt][1];
      }
    }
    if (ks == 0) kstep<NB8, true>(acc, ah, al, bh, bl);
    else kstep<NB8, false>(acc, ah, al, bh, bl);
  }
}

// ---- backward-data GEMM:  acc[in unit][batch] = sum_k W[unit][k] * Z[batch][k],  k = out unit
// standard k slots (t, t+4) and rows (g, g+8 <-> units 16w + g, 16w + g + 8); Z is stored with units
// k, k+4 adjacent.  W columns 50, 51 and Z units 50..55 are zero; k = 52..55 constant zeros.
template <int NB8, bool PRE_W, bool PRE_A>
__device__ __forceinline__ void gemm_bwd(const float* __restrict__ W, const float* __restrict__ Zb, int a_lo, int b_lo,
                                         float (&acc)[NB8][4], int w, int g, int t) {
  const float* a_base = W + (16 * w + g) * TW + t;
  const float* b_base = Zb + g * TH + 2 * t;
#pragma unroll
  for (int ks = 0; ks < 7; ++ks) {
    uint32_t ah[4], al[4], bh[NB8][2], bl[NB8][2];
    ld1<PRE_W>(a_base + 8 * ks, a_lo, ah[0], al[0]);
    ld1<PRE_W>(a_base + 8 * ks + 8 * TW, a_lo, ah[1], al[1]);
    if (ks < 6) {
      ld1<PRE_W>(a_base + 8 * ks + 4, a_lo, ah[2], al[2]);
      ld1<PRE_W>(a_base + 8 * ks + 4 + 8 * TW, a_lo, ah[3], al[3]);
    } else {
      ah[2] = al[2] = ah[3] = al[3] = 0u;
    }
#pragma unroll
    for (int nt = 0; nt < NB8; ++nt) {
      ld2<PRE_A>(b_base + 8 * nt * TH + 8 * ks, b_lo, bh[nt][0], bh[nt][1], bl[nt][0], bl[nt][1]);
      if (ks == 6) bh[nt][1] = bl[nt][1] = 0u;
    }
    if (ks == 0) kstep<NB8, true>(acc, ah, al, bh, bl);
    else kstep<NB8, false>(acc, ah, al, bh, bl);
  }
}

// ---- weight gradient of [W; b] of one hidden layer, straight to global memory:
//   grad[k][j] = pscale * Wb[k][j] + sum_i Hb[i][k] * Z[i][j]      k = in unit <= 50 (50: bias), j < 50
// M = in unit (rows g <-> 16w + 2g, g+8 <-> 16w + 2g + 1: one LDS.64 of Hb), N = out unit (7 n-tiles),
// K = batch with the standard slots (t, t+4); batch rows >= RB do not exist: constant zeros.
template <int NB8, int RB, bool PRE_W, bool PRE_A>
__device__ __forceinline__ void wgrad_t(const float* __restrict__ Hb, const float* __restrict__ Zb,
                                        const float* __restrict__ W, int w_lo, int a_lo, float* __restrict__ gW,
                                        float pscale, int w, int g, int t) {
  uint32_t ah[NB8][4], al[NB8][4];
  const float* a_base = Hb + t * TH + 16 * w + 2 * g;
#pragma unroll
  for (int ks = 0; ks < NB8; ++ks) {
    ld2<PRE_A>(a_base + 8 * ks * TH, a_lo, ah[ks][0], ah[ks][1], al[ks][0], al[ks][1]);
    if (8 * ks + 4 < RB) ld2<PRE_A>(a_base + (8 * ks + 4) * TH, a_lo, ah[ks][2], ah[ks][3], al[ks][2], al[ks][3]);
    else ah[ks][2] = al[ks][2] = ah[ks][3] = al[ks][3] = 0u;
  }
  const int k0 = 16 * w + 2 * g;                     // rows of c0, c1; c2, c3 belong to k0 + 1
  const float* b_base = Zb + t * TH + zpos(g);       // (zpos(8nt + g) = 8nt + zpos(g))
  // column tiles two at a time: 8 independent accumulator quads keep the tensor pipe fed
#pragma unroll
  for (int nt0 = 0; nt0 < NT8; nt0 += 2) {
    constexpr int NP = 2;
    float acc[NP][4];
#pragma unroll
    for (int ks = 0; ks < NB8; ++ks) {
      uint32_t bh[NP][2], bl[NP][2];
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int nt = nt0 + p < NT8 ? nt0 + p : NT8 - 1;
        ld1<PRE_A>(b_base + 8 * ks * TH + 8 * nt, a_lo, bh[p][0], bl[p][0]);
        if (8 * ks + 4 < RB) ld1<PRE_A>(b_base + (8 * ks + 4) * TH + 8 * nt, a_lo, bh[p][1], bl[p][1]);
        else bh[p][1] = bl[p][1] = 0u;
      }
      if (ks == 0) kstep<NP, true>(acc, ah[ks], al[ks], bh, bl);
      else kstep<NP, false>(acc, ah[ks], al[ks], bh, bl);
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int col = 8 * (nt0 + p) + 2 * t;
      if (nt0 + p < NT8 && col < HID) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int k = k0 + half;
          if (k <= HID) {
            float2 wv = *reinterpret_cast<const float2*>(W + k * TW + col);
            if constexpr (PRE_W) {
              const float2 wl = *reinterpret_cast<const float2*>(W + w_lo + k * TW + col);
              wv.x += wl.x; wv.y += wl.y;            // hi + lo is exact
            }
            *reinterpret_cast<float2*>(gW + k * HID + col) =
                make_float2(fmaf(wv.x, pscale, acc[p][2 * half]), fmaf(wv.y, pscale, acc[p][2 * half + 1]));
          }
        }
      }
    }
  }
}

// forward C layout (units 16w + 2g, +1; batch 8nt + 2t, +1) -> Hb[batch][unit] (pairs of units: 64-bit)
template <int NB8, int RB, bool PRE_A>
__device__ __forceinline__ void store_h(float* __restrict__ Hb, int lo_off, const float (&v)[NB8][4], int w, int g,
                                        int t) {
  const int unit = 16 * w + 2 * g;
  if (unit < HID) {
#pragma unroll
    for (int nt = 0; nt < NB8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int row = 8 * nt + 2 * t + e;
        if (row < RB) {
          float* q = Hb + row * TH + unit;
          if constexpr (PRE_A) {
            float h0, l0, h1, l1;
            split_rn(v[nt][e], h0, l0);
            split_rn(v[nt][2 + e], h1, l1);
            *reinterpret_cast<float2*>(q) = make_float2(h0, h1);
            *reinterpret_cast<float2*>(q + lo_off) = make_float2(l0, l1);
          } else {
            *reinterpret_cast<float2*>(q) = make_float2(v[nt][e], v[nt][2 + e]);
          }
        }
      }
  }
}

// dZ tile -> Z[batch][zpos(unit)].  FWD_LAYOUT: units 16w + 2g, +1 (head); else units 16w + g, +8.
template <int NB8, int RB, bool PRE_A, bool FWD_LAYOUT>
__device__ __forceinline__ void store_z(float* __restrict__ Zb, int lo_off, const float (&v)[NB8][4], int w, int g,
                                        int t) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int unit = FWD_LAYOUT ? 16 * w + 2 * g + half : 16 * w + g + 8 * half;
    if (unit < HID) {
#pragma unroll
      for (int nt = 0; nt < NB8; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int row = 8 * nt + 2 * t + e;
          if (row < RB) {
            float* q = Zb + row * TH + zpos(unit);
            if constexpr (PRE_A) {
              float h, l;
              split_rn(v[nt][2 * half + e], h, l);
              q[0] = h;
              q[lo_off] = l;
            } else {
              q[0] = v[nt][2 * half + e];
            }
          }
        }
    }
  }
}

// v <- v * (1 - h^2) in the backward C layout (units 16w + g, +8), h read back from Hb (hi + lo is exact)
template <int NB8, int RB, bool PRE_A>
__device__ __forceinline__ void times_tanh_prime(float (&v)[NB8][4], const float* __restrict__ Hb, int lo_off, int w,
                                                 int g, int t) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int unit = 16 * w + g + 8 * half;
#pragma unroll
    for (int nt = 0; nt < NB8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int row = 8 * nt + 2 * t + e;
        if (unit < HID && row < RB) {
          float h = Hb[row * TH + unit];
          if constexpr (PRE_A) h += Hb[lo_off + row * TH + unit];
          v[nt][2 * half + e] *= fmaf(-h, h, 1.0f);
        } else {
          v[nt][2 * half + e] = 0.0f;
        }
      }
  }
}

template <int NB8, int RB, bool WANT_GRAD, bool PRE_W, bool PRE_A>
__global__ void __launch_bounds__(T_THREADS, PRE_W ? 3 : 4) bnn_mma_t_kernel(BnnArgs a) {
  static_assert(RB <= 8 * NB8 && RB > 8 * NB8 - 8, "RB rows must end inside the last n-tile");
  extern __shared__ __align__(16) float smem_t[];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const BnnLayout L = a.L;
  const int batch = a.batch, n_in = L.n_in, D = L.D;
  constexpr int BP = 8 * NB8;                        // minibatch rows the n-tiles cover
  constexpr int WPL = TW * TW, APL = RB * TH;        // floats per weight / activation plane
  constexpr int W_LO = PRE_W ? WPL : 0, A_LO = PRE_A ? APL : 0;
  float* W2 = smem_t;
  float* W3 = W2 + (PRE_W ? 2 : 1) * WPL;
  float* H1 = W3 + (PRE_W ? 2 : 1) * WPL;
  float* H2 = H1 + (PRE_A ? 2 : 1) * APL;
  float* Zb = H2 + (PRE_A ? 2 : 1) * APL;
  float* fs = Zb + (PRE_A ? 2 : 1) * APL + 12 * TH;
  const int n_head = (n_in + 1) * HID;               // W1, b1
  float* sHead = fs;                                 // [n_head]
  float* sTail = sHead + (n_head + 3) / 4 * 4;       // W4[50], b4, rho (+ pad to 56)
  float* sX = sTail + 56;                            // [BP][n_in]
  float* sY = sX + BP * n_in;                        // [BP]
  float* scr = sY + BP;                              // [4][32] partial f | [16] sq partials
  // The padding every contraction runs over is written once: row 51 and columns 50, 51 of the weight
  // planes, units 50..55 of every activation row (unit 50 of H1 / H2: the 1 against the bias row).
  // Whatever else a fragment load can touch (rows of units >= 52, batch rows >= RB) only reaches
  // accumulator rows / columns that are discarded, or fragment words that are forced to zero.
  for (int q = tid; q < (PRE_W ? 4 : 2) * TW; q += T_THREADS) W2[(q / TW) * WPL + 51 * TW + q % TW] = 0.0f;
  for (int q = tid; q < (PRE_W ? 4 : 2) * 51 * 2; q += T_THREADS)
    W2[(q / 102) * WPL + ((q % 102) >> 1) * TW + HID + (q & 1)] = 0.0f;
  for (int q = tid; q < (PRE_A ? 6 : 3) * RB * 8; q += T_THREADS) {       // columns 48..55 (48, 49 are re-written)
    const int plane = q / (RB * 8), r = (q / 8) % RB, c = q % 8;
    const bool one = c == 2 && (PRE_A ? (plane == 0 || plane == 2) : plane < 2);
    H1[plane * APL + r * TH + 48 + c] = one ? 1.0f : 0.0f;
  }
  const float pscale = a.prior_den_inv * a.inv_n;
  const int uf = 16 * w + 2 * g;                     // forward layout: units uf, uf + 1
  const int ub = 16 * w + g;                         // backward layout: units ub, ub + 8

  for (int64_t chain = blockIdx.x; chain < a.n_chains; chain += gridDim.x) {
    const float* __restrict__ th = a.theta + chain * D;
    float* __restrict__ gr = WANT_GRAD ? a.grad + chain * D : nullptr;
    const int64_t start = a.starts != nullptr ? a.starts[chain] : 0;
    __syncthreads();                                 // the previous chain is done with every buffer
    // ---- stage theta: hidden layers into the W planes, the rest as fp32; sum of squares ----
    float sq = 0.0f;
    {
      const float2* src2 = reinterpret_cast<const float2*>(th + L.oW2);      // 8-byte aligned (D, oW2 even)
      constexpr int NPAIR = (HID + 1) * HID / 2;     // 1275 pairs per hidden layer, rows never straddled
#pragma unroll 1
      for (int q0 = 0; q0 < 2 * NPAIR; q0 += 5 * T_THREADS) {
        float2 v[5];
#pragma unroll
        for (int u = 0; u < 5; ++u) {
          const int q = q0 + u * T_THREADS + tid;
          v[u] = q < 2 * NPAIR ? __ldg(src2 + q) : make_float2(0.0f, 0.0f);
        }
#pragma unroll
        for (int u = 0; u < 5; ++u) {
          const int q = q0 + u * T_THREADS + tid;
          if (q < 2 * NPAIR) {
            const int layer = q >= NPAIR, e = 2 * (q - layer * NPAIR);
            const int k = e / HID, j = e - k * HID;
            float* dst = (layer ? W3 : W2) + k * TW + j;
            if constexpr (PRE_W) {
              float h0, l0, h1, l1;
              split_rn(v[u].x, h0, l0);
              split_rn(v[u].y, h1, l1);
              *reinterpret_cast<float2*>(dst) = make_float2(h0, h1);
              *reinterpret_cast<float2*>(dst + WPL) = make_float2(l0, l1);
            } else {
              *reinterpret_cast<float2*>(dst) = v[u];
            }
            sq = fmaf(v[u].x, v[u].x, sq);
            sq = fmaf(v[u].y, v[u].y, sq);
          }
        }
      }
      for (int q = tid; q < n_head; q += T_THREADS) {
        const float v = __ldg(th + q);
        sHead[q] = v;
        sq = fmaf(v, v, sq);
      }
      if (tid < HID + 2) {
        const float v = __ldg(th + L.oW4 + tid);
        sTail[tid] = v;
        sq = fmaf(v, v, sq);
      }
      for (int q = tid; q < BP * n_in; q += T_THREADS) sX[q] = q < batch * n_in ? __ldg(a.X + start * n_in + q) : 0.0f;
      if (tid < BP) sY[tid] = tid < batch ? __ldg(a.y + start + tid) : 0.0f;
      sq = warp_sum(sq);
      if (lane == 0) scr[128 + 4 + w] = sq;
    }
    __syncthreads();

    // ---- layer 1 (n_in -> 50) element-wise in the forward C layout, H1 <- tanh(X W1 + b1) ----
    float h[NB8][4], acc[NB8][4];
    {
      const float* W1 = sHead;
      const float* b1 = sHead + n_in * HID;
      const int c0 = min(uf, HID - 2);
      const float bb0 = b1[c0], bb1 = b1[c0 + 1];
#pragma unroll
      for (int nt = 0; nt < NB8; ++nt) { acc[nt][0] = acc[nt][1] = bb0; acc[nt][2] = acc[nt][3] = bb1; }
      for (int m = 0; m < n_in; ++m) {
        const float w0 = W1[m * HID + c0], w1 = W1[m * HID + c0 + 1];
#pragma unroll
        for (int nt = 0; nt < NB8; ++nt) {
          const float x0 = sX[(8 * nt + 2 * t) * n_in + m], x1 = sX[(8 * nt + 2 * t + 1) * n_in + m];
          acc[nt][0] = fmaf(x0, w0, acc[nt][0]); acc[nt][1] = fmaf(x1, w0, acc[nt][1]);
          acc[nt][2] = fmaf(x0, w1, acc[nt][2]); acc[nt][3] = fmaf(x1, w1, acc[nt][3]);
        }
      }
#pragma unroll
      for (int nt = 0; nt < NB8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) h[nt][e] = fast_tanh(acc[nt][e]);
      store_h<NB8, RB, PRE_A>(H1, A_LO, h, w, g, t);
    }
    __syncthreads();
    // ---- layers 2, 3 forward ----
    gemm_fwd<NB8, PRE_W, PRE_A>(W2, H1, W_LO, A_LO, acc, w, g, t);
#pragma unroll
    for (int nt = 0; nt < NB8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) h[nt][e] = fast_tanh(acc[nt][e]);
    store_h<NB8, RB, PRE_A>(H2, A_LO, h, w, g, t);
    __syncthreads();
    gemm_fwd<NB8, PRE_W, PRE_A>(W3, H2, W_LO, A_LO, acc, w, g, t);
    const bool live = uf < HID;                      // units uf, uf + 1 exist (HID is even)
#pragma unroll
    for (int nt = 0; nt < NB8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) h[nt][e] = live ? fast_tanh(acc[nt][e]) : 0.0f;     // H3, registers only

    // ---- head: f_i = sum_j H3[j][i] W4[j] + b4 ----
    const float w4_0 = live ? sTail[uf] : 0.0f, w4_1 = live ? sTail[uf + 1] : 0.0f;
#pragma unroll
    for (int nt = 0; nt < NB8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float f = fmaf(h[nt][e], w4_0, h[nt][2 + e] * w4_1);
        f += __shfl_xor_sync(0xffffffffu, f, 4);
        f += __shfl_xor_sync(0xffffffffu, f, 8);
        f += __shfl_xor_sync(0xffffffffu, f, 16);
        if (g == 0) scr[32 * w + 8 * nt + 2 * t + e] = f;
      }
    __syncthreads();
    const float b4 = sTail[HID], rho = sTail[HID + 1];
    const float e_rho = expf(rho);
    const float fvi = 1.0f / (e_rho + 1e-16f);                              // :368
    float df[NB8][2];
#pragma unroll
    for (int nt = 0; nt < NB8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = 8 * nt + 2 * t + e;
        const float f = ((scr[i] + scr[32 + i]) + (scr[64 + i] + scr[96 + i])) + b4;
        const float diff = i < batch ? sY[i] - f : 0.0f;
        df[nt][e] = -(diff * fvi) * a.inv_bs;                               // d cost / d f_i
      }
    if (w == 0) {                                    // cost (and the scalar gradients) by warp 0
      float diff = 0.0f;
      if (lane < batch) diff = sY[lane] - (((scr[lane] + scr[32 + lane]) + (scr[64 + lane] + scr[96 + lane])) + b4);
      const float sse_t = warp_sum(diff * diff);
      const float dfsum = warp_sum(-(diff * fvi) * a.inv_bs);
      if (lane == 0) {
        const float sq_t = (scr[132] + scr[133]) + (scr[134] + scr[135]);
        const float lv_den = 0.02f + 3e-16f;                                // safe_divide(., 2 * var)
        const float dl = rho - logf(1e-6f);
        const float log_like_data = (-sse_t * (0.5f * fvi) - 0.5f * rho * (float)batch) * a.inv_bs;
        const float lv = -(dl * dl) / lv_den - 0.5f * logf(0.01f);          // :102-107
        const float wp = (-0.5f * sq_t) * a.prior_den_inv;                  // :131-141
        a.cost[chain] = -(log_like_data + (lv + wp) * a.inv_n);
        if (a.mse != nullptr) a.mse[chain] = sse_t / (float)batch;
        if (WANT_GRAD) {
          const float drho_data = -(0.5f * sse_t * e_rho * fvi * fvi - 0.5f * (float)batch) * a.inv_bs;
          gr[L.orho] = drho_data + (2.0f * dl / lv_den) * a.inv_n + rho * pscale;
          gr[L.ob4] = fmaf(b4, pscale, dfsum);
        }
      }
    }
    if (!WANT_GRAD) continue;

    // ---- head backward: dW4[j] = sum_i H3[j][i] df_i;  dZ3 = (df W4^T) * (1 - H3^2) ----
    {
      float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
      for (int nt = 0; nt < NB8; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          s0 = fmaf(h[nt][e], df[nt][e], s0);
          s1 = fmaf(h[nt][2 + e], df[nt][e], s1);
        }
      s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      if (t == 0 && live) {
        gr[L.oW4 + uf] = fmaf(w4_0, pscale, s0);
        gr[L.oW4 + uf + 1] = fmaf(w4_1, pscale, s1);
      }
#pragma unroll
      for (int nt = 0; nt < NB8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v = h[nt][e];
          h[nt][e] = (df[nt][e & 1] * (e < 2 ? w4_0 : w4_1)) * fmaf(-v, v, 1.0f);
        }
    }
    store_z<NB8, RB, PRE_A, true>(Zb, A_LO, h, w, g, t);                    // dZ3
    __syncthreads();
    // ---- layer 3 backward ----
    gemm_bwd<NB8, PRE_W, PRE_A>(W3, Zb, W_LO, A_LO, acc, w, g, t);          // dH2 (units x batch)
    times_tanh_prime<NB8, RB, PRE_A>(acc, H2, A_LO, w, g, t);               // dZ2, registers
    wgrad_t<NB8, RB, PRE_W, PRE_A>(H2, Zb, W3, W_LO, A_LO, gr + L.oW3, pscale, w, g, t);
    __syncthreads();                                 // everyone has read dZ3
    store_z<NB8, RB, PRE_A, false>(Zb, A_LO, acc, w, g, t);                 // dZ2
    __syncthreads();
    // ---- layer 2 backward ----
    gemm_bwd<NB8, PRE_W, PRE_A>(W2, Zb, W_LO, A_LO, h, w, g, t);            // dH1
    times_tanh_prime<NB8, RB, PRE_A>(h, H1, A_LO, w, g, t);                 // dZ1, registers
    wgrad_t<NB8, RB, PRE_W, PRE_A>(H1, Zb, W2, W_LO, A_LO, gr + L.oW2, pscale, w, g, t);
    // ---- layer 1 backward from the registers (units ub, ub + 8):
    //      db1[j] = sum_i dZ1[j][i], dW1[m][j] = sum_i X[i][m] dZ1[j][i] ----
    {
      const int u0 = ub, u1 = ub + 8;
      float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
      for (int nt = 0; nt < NB8; ++nt) { s0 += h[nt][0] + h[nt][1]; s1 += h[nt][2] + h[nt][3]; }
      s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      const float* b1 = sHead + n_in * HID;
      if (t == 0) {
        if (u0 < HID) gr[L.ob1 + u0] = fmaf(b1[u0], pscale, s0);
        if (u1 < HID) gr[L.ob1 + u1] = fmaf(b1[u1], pscale, s1);
      }
      for (int m = 0; m < n_in; ++m) {
        float d0 = 0.0f, d1 = 0.0f;
#pragma unroll
        for (int nt = 0; nt < NB8; ++nt) {
          const float x0 = sX[(8 * nt + 2 * t) * n_in + m], x1 = sX[(8 * nt + 2 * t + 1) * n_in + m];
          d0 = fmaf(x0, h[nt][0], d0); d0 = fmaf(x1, h[nt][1], d0);
          d1 = fmaf(x0, h[nt][2], d1); d1 = fmaf(x1, h[nt][3], d1);
        }
        d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
        d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
        if (t == 0) {
          if (u0 < HID) gr[L.oW1 + m * HID + u0] = fmaf(sHead[m * HID + u0], pscale, d0);
          if (u1 < HID) gr[L.oW1 + m * HID + u1] = fmaf(sHead[m * HID + u1], pscale, d1);
        }
      }
    }
  }
}

}  // namespace sgmcmc
