// K4 on the tensor pipe: the BNN cost + gradient of one chain per CTA of NW warps (one warp
// per 16 minibatch rows) with mma.sync.m16n8k8 TF32 tiles and the 3xTF32 split
// (x = hi + lo, a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi; the dropped lo*lo term is ~2^-22
// relative), so results stay at fp32 accuracy (tests: same tolerances as the FFMA kernel).
// Replaces pysgmcmc/models/bayesian_neural_network.py:28-69,77-141,337-388 + tf.gradients.
//
// Why this shape.  The FFMA kernel (bnn.cu) is bound by shared-memory operand loads
// (one LDS.128 per 4 FFMA per thread, ~2.6 cycles each that do not overlap FFMA issue).
// An mma.sync tile reads each operand word once per 16x8x8 block, and the legacy tensor
// path of sm_100a delivers 277 TFLOP/s TF32 (tools/micro/mma_tf32_bench.cu, 8.6 cycles per
// m16n8k8 per SM sub-partition): 3 split products at 56 % tile occupancy still leave ~3.5x
// the FFMA kernel's rate.  tcgen05 would need M >= 64 per chain-private weight matrix and
// operands re-laid out in shared memory per chain; see DESIGN.md "K4-MMA".
//
// Data flow (lane = 4*g + t, warp w owns minibatch rows 16w .. 16w+15):
//  * the chain's whole parameter row theta[D] is staged ONCE in shared memory in its natural
//    layout (R).  Because b_l follows W_l in that layout, [W_l; b_l] is a 51 x 50 matrix with
//    row stride 50: the bias add (forward) and the bias gradient (backward) come out of the
//    same MMAs by giving every activation matrix a constant-1 column 50.
//  * activations live in REGISTERS in the mma C layout (rows g, g+8; columns 2t, 2t+1 of
//    every 8-wide tile).  The k index of the next GEMM is a summation index, so a C tile is
//    re-used directly as an A fragment (k slot t <-> column 2t, slot t+4 <-> column 2t+1) and
//    the B fragment of the weights is gathered with the same permutation: the forward chain
//    H1 -> H2 -> H3 and the backward chain dZ3 -> dZ2 -> dZ1 never leave the register file,
//    and the warps of a chain do not exchange anything in those phases.
//  * only the weight-gradient GEMMs dW = [H 1]^T dZ contract over the batch index, which is
//    spread over lanes and warps: H1, H2 (forward) and the current dZ are also written to
//    three small [batch x 56] shared buffers and re-read as transposed fragments
//    (conflict-free: stride 56); the 4 row tiles of dW are split between the warps.
//  * the gradient overwrites R in place (each dW element is produced by the thread that read
//    the weight for the prior term) and leaves with coalesced 128-bit stores -- or, in the
//    fused K5 kernel, feeds the SGHMC update without touching HBM.
#pragma once

#include "bnn_common.cuh"

namespace sgmcmc {

constexpr int AS = 56;   // row stride of the activation buffers: 50 units, the 1-column, 5 zeros
constexpr int NT8 = 7;   // 8-wide tiles covering those 56 columns

// x = hi + lo with hi = x truncated to TF32 -- which is what the tensor core does to a raw
// fp32 operand (it reads the top 19 bits), so hi costs no instruction -- and lo the exact fp32
// remainder (2 instructions).  |lo| < 2^-10 |x|; the hardware truncates lo to 11 bits too.
// Measured on the B200 against the float64 oracle: gradient error 4e-8 rms / 1e-6 max of
// max|g| per chain (the FFMA kernel: 1e-8 / 4e-7).  SGMCMC_TF32_ROUND_SPLIT=1 rounds hi instead
// (one more integer add per split, unbiased lo): an emulation of 51-term dot products shows
// 3.7x less split error, but on the device the total error does not move (it is dominated by
// the tensor core's own fp32 accumulation) while K4 slows from 0.218 to 0.223 ms, so the
// truncating split is the default.
//
// Accuracy modes (template parameter MODE of everything below; launch variants 10-13):
//   bit 0 (MMA_ROUND_SPLIT): hi is rounded to nearest instead of truncated, so lo is symmetric
//         around 0 and the dropped lo*lo term (2^-22 of every product with the truncating
//         split, always of the product's sign) stops being a systematic bias;
//   bit 1 (MMA_RN_ACCUM): the three products of a k-step go into a zeroed register tile and are
//         added to the running sum by the FP32 pipe (round to nearest) -- the tensor core adds
//         into its accumulator operand with truncation, a bias that compounds over the k-steps
//         of a GEMM (Ootomo & Yokota 2022).
// What they buy is measured as trajectory drift, not single-step error
// (tools/bnn_trajectory_drift.py, profiles/r02_bnn_trajectory_drift*.jsonl).
constexpr int MMA_ROUND_SPLIT = 1, MMA_RN_ACCUM = 2, MMA_PACKED_SPLIT = 4, MMA_SEP_CROSS = 8;
// the default (launch variant 16): every fused / pipelined caller of bnn_chain_mma uses it, so that all of
// them stay bit-identical to K4 then K1
constexpr int MMA_DEFAULT_MODE = MMA_ROUND_SPLIT | MMA_RN_ACCUM | MMA_SEP_CROSS | MMA_PACKED_SPLIT;
template <int MODE>
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  if constexpr ((MODE & MMA_ROUND_SPLIT) != 0) hi = __float_as_uint(x) + 0x1000u;
  else hi = __float_as_uint(x);
  lo = __float_as_uint(x - __uint_as_float(hi & 0xffffe000u));
}

// Two operand words at once.  MMA_PACKED_SPLIT (launch variant 14 = 13 + this, for the weight fragments):
// Veltkamp's splitting on the FP32 pipe, packed
//   p = 8193 x;  hi = p - 8192 x  (x rounded to 11 significant bits, exactly a TF32 value);  lo = x - hi
// = FMUL2 + FFMA2 + FADD2 for the pair, 1.5 issue slots per word instead of 3 (integer add, mask,
// subtract): the hi/lo split is 32 % of the instructions K4 executes (profiles/r02_ncu_k4_mode13_summary.txt)
// and the kernel is bound by issue slots, not by the FP32 pipe the packed instructions occupy for two cycles.
// Measured: 0.255 against 0.266 ms -- and a 1000-step trajectory that ends just ABOVE 1e-5 where the
// integer split ends just below it (both are roundings to nearest; the difference is the trajectory's
// own sensitivity), so it is not the default.
template <int MODE>
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& h0, uint32_t& h1, uint32_t& l0, uint32_t& l1) {
  if constexpr ((MODE & MMA_PACKED_SPLIT) != 0) {
    asm("{\n\t.reg .b64 x, p, h, l, c1, c2;\n\t"
        "mov.b64 x, {%4, %5};\n\t"
        "mov.b64 c1, {%6, %6};\n\t"
        "mov.b64 c2, {%7, %7};\n\t"
        "mul.rn.f32x2 p, x, c1;\n\t"
        "fma.rn.f32x2 h, x, c2, p;\n\t"
        "sub.rn.f32x2 l, x, h;\n\t"
        "mov.b64 {%0, %1}, h;\n\t"
        "mov.b64 {%2, %3}, l;\n\t}"
        : "=r"(h0), "=r"(h1), "=r"(l0), "=r"(l1)
        : "f"(x0), "f"(x1), "f"(8193.0f), "f"(-8192.0f));
  } else {
    split_tf32<MODE>(x0, h0, l0);
    split_tf32<MODE>(x1, h1, l1);
  }
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// d = a * b (a zero accumulator operand: the register-zero form, no tile to clear first)
__device__ __forceinline__ void mma_tf32_zero(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "f"(0.0f));
}

// acc(2) += part(2), round to nearest, one packed instruction (FADD2)
__device__ __forceinline__ void add2_rn(float& a0, float& a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tadd.rn.f32x2 ra, ra, rb;\n\t"
      "mov.b64 {%0, %1}, ra;\n\t}"
      : "+f"(a0), "+f"(a1)
      : "f"(b0), "f"(b1));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Barrier among the NW warps that work on one chain.  bar_id == 0: they are the whole CTA; bar_id > 0: they
// are one group of a warp-specialised CTA (bnn_fused.cu) and meet at that named barrier.
template <int NW>
__device__ __forceinline__ void chain_barrier(int bar_id = 0) {
  if constexpr (NW == 1) {
    __syncwarp();
  } else {
    if (bar_id == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * NW) : "memory");
  }
}

constexpr int K4_PREFETCH_AHEAD = 1024;   // chains between a CTA and the row it prefetches into L2 (0: off)
constexpr int MMA_SCRATCH = 192;   // cross-warp partial sums: [2][64] dW4 columns, [64] scalars
// shared memory per chain (floats): R[D rounded to 4] | P | Q | Z ([batch x AS] each) | X | y | scratch
__host__ __device__ inline int bnn_mma_smem_floats(int batch, int n_in, int D) {
  return ((D + 3) & ~3) + 3 * batch * AS + ((batch * n_in + 3) & ~3) + ((batch + 3) & ~3) + MMA_SCRATCH;
}

// C-layout register tile (rows r0 = 16w+g, r0+8) <-> [batch x AS] shared buffer (64-bit accesses;
// a half warp touches rows g = 0..3 at bank offsets 0, 24, 16, 8: conflict-free)
__device__ __forceinline__ void store_c(float* __restrict__ buf, int batch, int r0, int t,
                                        const float (&h)[NT8][4]) {
  const int r1 = r0 + 8;
#pragma unroll
  for (int nt = 0; nt < NT8; ++nt) {
    if (r0 < batch) *reinterpret_cast<float2*>(buf + r0 * AS + 8 * nt + 2 * t) = make_float2(h[nt][0], h[nt][1]);
    if (r1 < batch) *reinterpret_cast<float2*>(buf + r1 * AS + 8 * nt + 2 * t) = make_float2(h[nt][2], h[nt][3]);
  }
}

// A fragments (hi, lo) of k-step ks taken from the C-layout tile ks of `src`
template <int MODE>
__device__ __forceinline__ void a_from_c(const float (&src)[NT8][4], int ks, uint32_t (&ah)[4], uint32_t (&al)[4]) {
  // (row g, k slot t) = column 2t, (row g+8, k slot t) | (row g, k slot t+4) = column 2t+1, (row g+8, k slot t+4)
  split_tf32<MODE>(src[ks][0], ah[0], al[0]);
  split_tf32<MODE>(src[ks][2], ah[1], al[1]);
  split_tf32<MODE>(src[ks][1], ah[2], al[2]);
  split_tf32<MODE>(src[ks][3], ah[3], al[3]);
}

//   bit 3 (MMA_SEP_CROSS, with MMA_RN_ACCUM): the two cross terms of every k-step are chained through the
//         tensor core into their OWN accumulator `corr` -- their sum is 2^-11 of the result, so is the ulp that
//         the tensor core truncates at -- and only the exact hi*hi products of a k-step go through a zero C
//         operand and the FP32-pipe addition; `corr` is added once at the end of the GEMM.  Mode 13 adds the
//         cross terms and the hi*hi products inside the tensor core, one truncation at the ulp of the k-step's
//         sum per k-step; this removes it.
// FIRST: the first k-step of a GEMM -- the products go straight into the (not yet initialised) accumulators
// through the zero-C form (0 + x = x exactly: the same bits as clearing the tile and adding)
template <int MODE, bool FIRST = false>
__device__ __forceinline__ void mma3_row(float (&acc)[NT8][4], float (&corr)[NT8][4], const uint32_t (&ah)[4],
                                         const uint32_t (&al)[4], const uint32_t (&bh)[NT8][2],
                                         const uint32_t (&bl)[NT8][2]) {
  // three passes over 7 independent accumulators: dependent MMAs are 7 issues apart
  if constexpr ((MODE & MMA_SEP_CROSS) != 0) {
    if constexpr (FIRST) {
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) mma_tf32_zero(corr[nt], al, bh[nt]);
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) mma_tf32_zero(acc[nt], ah, bh[nt]);
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) mma_tf32(corr[nt], ah, bl[nt]);
    } else {
      float part[NT8][4];
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) mma_tf32(corr[nt], al, bh[nt]);
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) mma_tf32_zero(part[nt], ah, bh[nt]);
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) mma_tf32(corr[nt], ah, bl[nt]);
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) {
        add2_rn(acc[nt][0], acc[nt][1], part[nt][0], part[nt][1]);
        add2_rn(acc[nt][2], acc[nt][3], part[nt][2], part[nt][3]);
      }
    }
  } else if constexpr (FIRST) {
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) mma_tf32_zero(acc[nt], al, bh[nt]);
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) mma_tf32(acc[nt], ah, bl[nt]);
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) mma_tf32(acc[nt], ah, bh[nt]);
  } else if constexpr ((MODE & MMA_RN_ACCUM) != 0) {
    float part[NT8][4];
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) mma_tf32_zero(part[nt], al, bh[nt]);
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) mma_tf32(part[nt], ah, bl[nt]);
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) mma_tf32(part[nt], ah, bh[nt]);
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) {               // packed add.rn.f32x2: two accumulator elements per issue slot
      add2_rn(acc[nt][0], acc[nt][1], part[nt][0], part[nt][1]);
      add2_rn(acc[nt][2], acc[nt][3], part[nt][2], part[nt][3]);
    }
  } else {
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) mma_tf32(acc[nt], al, bh[nt]);
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) mma_tf32(acc[nt], ah, bl[nt]);
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) mma_tf32(acc[nt], ah, bh[nt]);
  }
}

// acc += corr at the end of a GEMM (MMA_SEP_CROSS)
template <int MODE>
__device__ __forceinline__ void add_cross(float (&acc)[NT8][4], const float (&corr)[NT8][4]) {
  if constexpr ((MODE & MMA_SEP_CROSS) != 0) {
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) {
      add2_rn(acc[nt][0], acc[nt][1], corr[nt][0], corr[nt][1]);
      add2_rn(acc[nt][2], acc[nt][3], corr[nt][2], corr[nt][3]);
    }
  }
}

__device__ __forceinline__ void zero_tile(float (&acc)[NT8][4]) {
#pragma unroll
  for (int nt = 0; nt < NT8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f;
}

// acc[i][j] = sum_{k<=50} src[i][k] * Wb[k][j]   (Wb = [W; b], row stride 50; src column 50 is 1)
template <int MODE>
__device__ __forceinline__ void gemm_forward(const float* __restrict__ Wb, const float (&src)[NT8][4],
                                             float (&acc)[NT8][4], int g, int t) {
  float corr[NT8][4];                                    // (MMA_SEP_CROSS only; dead otherwise)
#pragma unroll
  for (int ks = 0; ks < NT8; ++ks) {
    uint32_t ah[4], al[4], bh[NT8][2], bl[NT8][2];
    a_from_c<MODE>(src, ks, ah, al);
    const int k0 = 8 * ks + 2 * t, k1 = k0 + 1;          // rows of [W; b] behind k slots t, t+4
    const float* w0 = Wb + k0 * HID + g;
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) {
      // columns 50..55 of the last tile read the head of the next row: finite, and their
      // outputs are masked by the caller.  Rows > 50 do not exist.
      float b0, b1;
      if (ks < NT8 - 1) {
        b0 = w0[8 * nt];
        b1 = w0[HID + 8 * nt];
      } else {                                             // rows 48 + 2t (+1): clamp, then mask
        b0 = Wb[min(k0, HID) * HID + g + 8 * nt];
        b1 = Wb[min(k1, HID) * HID + g + 8 * nt];
        b0 = k0 <= HID ? b0 : 0.0f;
        b1 = k1 <= HID ? b1 : 0.0f;
      }
      split_pair<MODE>(b0, b1, bh[nt][0], bh[nt][1], bl[nt][0], bl[nt][1]);
    }
    if (ks == 0) mma3_row<MODE, true>(acc, corr, ah, al, bh, bl);
    else mma3_row<MODE>(acc, corr, ah, al, bh, bl);
  }
  add_cross<MODE>(acc, corr);
}

// acc[i][k] = sum_j src[i][j] * W[k][j]   (src columns >= 50 are 0)
template <int MODE>
__device__ __forceinline__ void gemm_backward_data(const float* __restrict__ Wb, const float (&src)[NT8][4],
                                                   float (&acc)[NT8][4], int g, int t) {
  float corr[NT8][4];                                    // (MMA_SEP_CROSS only; dead otherwise)
#pragma unroll
  for (int ks = 0; ks < NT8; ++ks) {
    uint32_t ah[4], al[4], bh[NT8][2], bl[NT8][2];
    a_from_c<MODE>(src, ks, ah, al);
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) {
      const int k = min(8 * nt + g, HID - 1);            // output units >= 50 are discarded
      const float2 b = *reinterpret_cast<const float2*>(Wb + k * HID + 8 * ks + 2 * t);
      split_pair<MODE>(b.x, b.y, bh[nt][0], bh[nt][1], bl[nt][0], bl[nt][1]);
    }
    if (ks == 0) mma3_row<MODE, true>(acc, corr, ah, al, bh, bl);
    else mma3_row<MODE>(acc, corr, ah, al, bh, bl);
  }
  add_cross<MODE>(acc, corr);
}

// Wb[k][j] <- pscale * Wb[k][j] + sum_i Hb[i][k] * Zb[i][j]   for k <= 50 (row 50: bias), j < 50
// i.e. the gradient of [W; b] (weight prior included) replaces the weights in place.  The
// MTW row tiles (16 rows k each) from mt0 on are this warp's share; their A fragments stay
// in registers while the column tiles are walked two at a time (4 independent accumulators).
template <int NB8, int MTW, int MODE>
__device__ __forceinline__ void gemm_weight_grad(const float* __restrict__ Hb, const float* __restrict__ Zb,
                                                 float* __restrict__ Wb, int batch, float pscale, int mt0,
                                                 int g, int t) {
  uint32_t ah[MTW][NB8][4], al[MTW][NB8][4];
#pragma unroll
  for (int m = 0; m < MTW; ++m) {
    const int k0 = 16 * (mt0 + m) + g, k1 = k0 + 8;
#pragma unroll
    for (int ks = 0; ks < NB8; ++ks) {
      const int i0 = 8 * ks + t, i1 = i0 + 4;
      // clamped addresses + selects: no branches around the loads
      const float* h0 = Hb + min(i0, batch - 1) * AS;
      const float* h1 = Hb + min(i1, batch - 1) * AS;
      const int k1c = min(k1, AS - 1);
      float a0 = h0[k0], a1 = h0[k1c], a2 = h1[k0], a3 = h1[k1c];
      a0 = i0 < batch ? a0 : 0.0f;
      a1 = (i0 < batch && k1 < AS) ? a1 : 0.0f;
      a2 = i1 < batch ? a2 : 0.0f;
      a3 = (i1 < batch && k1 < AS) ? a3 : 0.0f;
      split_pair<MODE>(a0, a1, ah[m][ks][0], ah[m][ks][1], al[m][ks][0], al[m][ks][1]);
      split_pair<MODE>(a2, a3, ah[m][ks][2], ah[m][ks][3], al[m][ks][2], al[m][ks][3]);
    }
  }
#pragma unroll
  for (int nt0 = 0; nt0 < NT8; nt0 += 2) {
    constexpr int NP = 2;
    uint32_t bh[NP][NB8][2], bl[NP][NB8][2];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int nt = nt0 + p < NT8 ? nt0 + p : NT8 - 1;
#pragma unroll
      for (int ks = 0; ks < NB8; ++ks) {
        const int i0 = 8 * ks + t, i1 = i0 + 4;
        float b0 = Zb[min(i0, batch - 1) * AS + 8 * nt + g];
        float b1 = Zb[min(i1, batch - 1) * AS + 8 * nt + g];
        b0 = i0 < batch ? b0 : 0.0f;
        b1 = i1 < batch ? b1 : 0.0f;
        split_pair<MODE>(b0, b1, bh[p][ks][0], bh[p][ks][1], bl[p][ks][0], bl[p][ks][1]);
      }
    }
    float acc[MTW][NP][4], corr[MTW][NP][4];
    constexpr bool SEP = (MODE & MMA_SEP_CROSS) != 0;
#pragma unroll
    for (int ks = 0; ks < NB8; ++ks) {
      float part[MTW][NP][4];
      // FP32-pipe accumulation: every k-step after the first goes into `part` (zero C operand) and is added
      // RN; the first one -- and all of them in the chained modes -- goes straight into the accumulators.
      // SEP: the cross terms chain into `corr`, only hi*hi takes that route.
      const bool RN = (MODE & MMA_RN_ACCUM) != 0 && ks > 0;        // (compile time: the loop is unrolled)
      auto& dst = *(RN ? &part : &acc);
      auto& small = *(SEP ? &corr : &dst);
#pragma unroll
      for (int m = 0; m < MTW; ++m)
#pragma unroll
        for (int p = 0; p < NP; ++p)
          if (nt0 + p < NT8) {
            if ((SEP ? ks == 0 : (RN || ks == 0))) mma_tf32_zero(small[m][p], al[m][ks], bh[p][ks]);
            else mma_tf32(small[m][p], al[m][ks], bh[p][ks]);
          }
      if (SEP) {
#pragma unroll
        for (int m = 0; m < MTW; ++m)
#pragma unroll
          for (int p = 0; p < NP; ++p)
            if (nt0 + p < NT8) mma_tf32_zero(dst[m][p], ah[m][ks], bh[p][ks]);
#pragma unroll
        for (int m = 0; m < MTW; ++m)
#pragma unroll
          for (int p = 0; p < NP; ++p)
            if (nt0 + p < NT8) mma_tf32(small[m][p], ah[m][ks], bl[p][ks]);
      } else {
#pragma unroll
        for (int m = 0; m < MTW; ++m)
#pragma unroll
          for (int p = 0; p < NP; ++p)
            if (nt0 + p < NT8) mma_tf32(dst[m][p], ah[m][ks], bl[p][ks]);
#pragma unroll
        for (int m = 0; m < MTW; ++m)
#pragma unroll
          for (int p = 0; p < NP; ++p)
            if (nt0 + p < NT8) mma_tf32(dst[m][p], ah[m][ks], bh[p][ks]);
      }
      if (RN) {
#pragma unroll
        for (int m = 0; m < MTW; ++m)
#pragma unroll
          for (int p = 0; p < NP; ++p)
            if (nt0 + p < NT8) {
              add2_rn(acc[m][p][0], acc[m][p][1], part[m][p][0], part[m][p][1]);
              add2_rn(acc[m][p][2], acc[m][p][3], part[m][p][2], part[m][p][3]);
            }
      }
    }
    if (SEP) {
#pragma unroll
      for (int m = 0; m < MTW; ++m)
#pragma unroll
        for (int p = 0; p < NP; ++p)
          if (nt0 + p < NT8) {
            add2_rn(acc[m][p][0], acc[m][p][1], corr[m][p][0], corr[m][p][1]);
            add2_rn(acc[m][p][2], acc[m][p][3], corr[m][p][2], corr[m][p][3]);
          }
    }
#pragma unroll
    for (int m = 0; m < MTW; ++m) {
      const int k0 = 16 * (mt0 + m) + g, k1 = k0 + 8;
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int col = 8 * (nt0 + p) + 2 * t;
        if (nt0 + p < NT8 && col < HID) {
          if (k0 <= HID) {
            float2* q = reinterpret_cast<float2*>(Wb + k0 * HID + col);
            const float2 wv = *q;
            *q = make_float2(fmaf(wv.x, pscale, acc[m][p][0]), fmaf(wv.y, pscale, acc[m][p][1]));
          }
          if (k1 <= HID) {
            float2* q = reinterpret_cast<float2*>(Wb + k1 * HID + col);
            const float2 wv = *q;
            *q = make_float2(fmaf(wv.x, pscale, acc[m][p][2]), fmaf(wv.y, pscale, acc[m][p][3]));
          }
        }
      }
    }
  }
}

struct BnnMmaSmem {
  float *R, *P, *Q, *Zb, *sX, *sY, *scr;
};

__device__ __forceinline__ BnnMmaSmem bnn_mma_carve(float* base, int batch, int n_in, int D) {
  BnnMmaSmem s;
  s.R = base;
  s.P = s.R + ((D + 3) & ~3);
  s.Q = s.P + batch * AS;
  s.Zb = s.Q + batch * AS;
  s.sX = s.Zb + batch * AS;
  s.sY = s.sX + ((batch * n_in + 3) & ~3);
  s.scr = s.sY + ((batch + 3) & ~3);
  return s;
}

// The cost and (WANT_GRAD) the gradient of ONE chain, by the NW = ceil(NB8 / 2) warps of a
// CTA.  On return (after a chain barrier) s.R holds the gradient in the parameter layout
// (theta is gone) and thread 0 has the cost and the sum of squared errors.
// `th` is the chain's parameter row in global memory (staged here into R).  COHERENT: the
// calling kernel also WRITES theta (K5), so the row is read with ld.global.cg (L2, where the
// update phase re-reads it) instead of the read-only path.
// BATCH_CT > 0: the minibatch has exactly that many rows (the launcher checked), so every row mask of the
// fragment loads folds at compile time -- for the reference's 20 rows (16 + 4: the k slots t / t + 4 of the
// third k-step are all live / all dead) that is every FSEL and most ISETPs of the weight-gradient GEMMs.
template <int NB8, bool WANT_GRAD, bool COHERENT = false, int MODE = 0, int BATCH_CT = 0>
__device__ __forceinline__ void bnn_chain_mma(const BnnArgs& a, const float* __restrict__ th,
                                              const int32_t* __restrict__ start_ptr, const BnnMmaSmem& s,
                                              float& cost_out, float& sse_out, int tid0 = 0, int bar_id = 0) {
  constexpr int NW = (NB8 + 1) / 2;
  constexpr int NTHR = 32 * NW;
  const int tid = threadIdx.x - tid0, lane = tid & 31, w = tid >> 5;     // (tid0: first thread of this chain's group)
  const int g = lane >> 2, t = lane & 3;
  const BnnLayout L = a.L;
  const int batch = BATCH_CT > 0 ? BATCH_CT : a.batch, n_in = L.n_in, D = L.D;
  float* __restrict__ R = s.R;

  // ---- stage theta (natural layout) and the minibatch; sum of squares for the weight prior ----
  float sq = 0.0f;
  const int64_t start = start_ptr != nullptr ? *start_ptr : 0;
  if ((D & 3) == 0 && aligned_to_dev(th, 16)) {
    const float4* src = reinterpret_cast<const float4*>(th);
    float4* dst = reinterpret_cast<float4*>(R);
    const int n4 = D / 4;
    // STAGE_U independent 128-bit loads per thread in flight: the whole row of the default network (1313
    // float4 over 64 threads) in ONE round -- the registers are free at this point, and every further round
    // exposes another global-memory latency at the head of the chain (10 % of K4's stall samples with 8)
    constexpr int STAGE_U = NW == 2 ? 21 : 16;
    for (int q0 = 0; q0 < n4; q0 += STAGE_U * NTHR) {
      float4 v[STAGE_U];
#pragma unroll
      for (int u = 0; u < STAGE_U; ++u) {
        const int q = q0 + u * NTHR + tid;
        v[u] = q < n4 ? (COHERENT ? __ldcg(src + q) : __ldg(src + q)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      }
#pragma unroll
      for (int u = 0; u < STAGE_U; ++u) {
        const int q = q0 + u * NTHR + tid;
        if (q < n4) dst[q] = v[u];
        sq = fmaf(v[u].x, v[u].x, sq); sq = fmaf(v[u].y, v[u].y, sq);
        sq = fmaf(v[u].z, v[u].z, sq); sq = fmaf(v[u].w, v[u].w, sq);
      }
    }
  } else {
#pragma unroll 4
    for (int q = tid; q < D; q += NTHR) {
      const float v = COHERENT ? __ldcg(th + q) : __ldg(th + q);
      R[q] = v;
      sq = fmaf(v, v, sq);
    }
  }
  for (int q = tid; q < batch * n_in; q += NTHR) s.sX[q] = __ldg(a.X + start * n_in + q);
  for (int q = tid; q < batch; q += NTHR) s.sY[q] = __ldg(a.y + start + q);
  sq = warp_sum(sq);
  chain_barrier<NW>(bar_id);

  // ---- layer 1 (n_in -> 50), element-wise in the C layout: h <- [tanh(X W1 + b1) 1 0..] ----
  const int r0 = 16 * w + g, r1 = r0 + 8;
  float h[NT8][4], acc[NT8][4];
#pragma unroll
  for (int nt = 0; nt < NT8; ++nt) {
    const int col = 8 * nt + 2 * t;
    const float2 b = *reinterpret_cast<const float2*>(R + L.ob1 + min(col, HID - 2));
    acc[nt][0] = acc[nt][2] = b.x;
    acc[nt][1] = acc[nt][3] = b.y;
  }
  for (int m = 0; m < n_in; ++m) {
    const float x0 = r0 < batch ? s.sX[r0 * n_in + m] : 0.0f;
    const float x1 = r1 < batch ? s.sX[r1 * n_in + m] : 0.0f;
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) {
      const int col = 8 * nt + 2 * t;
      const float2 wv = *reinterpret_cast<const float2*>(R + L.oW1 + m * HID + min(col, HID - 2));
      acc[nt][0] = fmaf(x0, wv.x, acc[nt][0]); acc[nt][1] = fmaf(x0, wv.y, acc[nt][1]);
      acc[nt][2] = fmaf(x1, wv.x, acc[nt][2]); acc[nt][3] = fmaf(x1, wv.y, acc[nt][3]);
    }
  }
  // activation of a C tile: columns 48..55 of the last tile are units 48, 49, the 1, zeros;
  // row groups beyond the minibatch are skipped (their values are never used)
  const bool live1 = 16 * w + 8 < 8 * NB8;               // warp-uniform: rows r1 exist at all
  auto activate = [&](void) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      if (half == 0 || live1) {
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt) {
          float v[2];
          fast_tanh2(acc[nt][2 * half], acc[nt][2 * half + 1], v[0], v[1]);
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (nt == NT8 - 1) {
              const int c = 2 * t + e;
              v[e] = c < 2 ? v[e] : (c == 2 ? 1.0f : 0.0f);
            }
            h[nt][2 * half + e] = v[e];
          }
        }
      } else {
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt) h[nt][2] = h[nt][3] = 0.0f;
      }
    }
  };
  activate();
  if (WANT_GRAD) store_c(s.P, batch, r0, t, h);

  // ---- layers 2, 3 forward: h <- tanh([h 1] [W; b]) ----
#pragma unroll 1
  for (int l = 0; l < 2; ++l) {
    gemm_forward<MODE>(R + (l == 0 ? L.oW2 : L.oW3), h, acc, g, t);
    activate();
    if (l == 0 && WANT_GRAD) store_c(s.Q, batch, r0, t, h);
  }

  // ---- head: f_i = [H3 1][i, :] . [W4; b4]; loss pieces (bayesian_neural_network.py:368-388) ----
  float w4[NT8][2];
#pragma unroll
  for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = 8 * nt + 2 * t + e;
      w4[nt][e] = col <= HID ? R[L.oW4 + col] : 0.0f;     // col 50 is b4 (it follows W4)
    }
  const float rho = R[L.orho];
  const float e_rho = expf(rho);
  const float fvi = 1.0f / (e_rho + 1e-16f);                        // :368
  float df[2], sse = 0.0f;
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    float f = 0.0f;
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) {
      f = fmaf(h[nt][2 * rr], w4[nt][0], f);
      f = fmaf(h[nt][2 * rr + 1], w4[nt][1], f);
    }
    f += __shfl_xor_sync(0xffffffffu, f, 1);
    f += __shfl_xor_sync(0xffffffffu, f, 2);
    const int row = r0 + 8 * rr;
    const float diff = row < batch ? s.sY[row] - f : 0.0f;
    df[rr] = -(diff * fvi) * a.inv_bs;                              // d cost / d f_i
    if (t == 0) sse = fmaf(diff, diff, sse);
  }
  sse = warp_sum(sse);
  float* scr = s.scr;
  if (lane == 0) { scr[128 + w] = sse; scr[136 + w] = sq; }
  const float pscale = a.prior_den_inv * a.inv_n;

  if (WANT_GRAD) {
    // ---- layer 4 backward: dW4 (and db4 in column 50) = sum_i [H3 1][i][:] df_i; then
    //      h <- dZ3 = (df W4^T) * (1 - H3^2): the 1-column gives 0, columns > 50 have w4 = 0 ----
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float sum = fmaf(h[nt][e], df[0], h[nt][2 + e] * df[1]);
        sum += __shfl_xor_sync(0xffffffffu, sum, 4);
        sum += __shfl_xor_sync(0xffffffffu, sum, 8);
        sum += __shfl_xor_sync(0xffffffffu, sum, 16);
        if (g == 0) scr[64 * w + 8 * nt + 2 * t + e] = sum;
      }
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float v = h[nt][e];
        const float wv = (nt == NT8 - 1 && 2 * t + (e & 1) >= 2) ? 0.0f : w4[nt][e & 1];
        h[nt][e] = (df[e >> 1] * wv) * fmaf(-v, v, 1.0f);
      }
  }
  chain_barrier<NW>(bar_id);                 // every thread has read W4, b4, rho; partial sums are visible
  if (tid == 0) {
    float sse_t = 0.0f, sq_t = 0.0f;
#pragma unroll
    for (int ww = 0; ww < NW; ++ww) { sse_t += scr[128 + ww]; sq_t += scr[136 + ww]; }
    const float lv_den = 0.02f + 3e-16f;                            // safe_divide(., 2 * var)
    const float dl = rho - logf(1e-6f);
    // un-contracted on purpose: which product the compiler fuses into an FMA depends on what it knows at
    // compile time (BATCH_CT), and every instantiation of this function has to produce the same bits
    const float nb = (float)batch;
    const float log_like_data = __fmul_rn(__fsub_rn(__fmul_rn(-sse_t, __fmul_rn(0.5f, fvi)),
                                                    __fmul_rn(__fmul_rn(0.5f, rho), nb)), a.inv_bs);
    const float lv = __fsub_rn(__fdiv_rn(-__fmul_rn(dl, dl), lv_den), 0.5f * logf(0.01f));      // :102-107
    const float wp = __fmul_rn(__fmul_rn(-0.5f, sq_t), a.prior_den_inv);                         // :131-141
    cost_out = -__fadd_rn(log_like_data, __fmul_rn(__fadd_rn(lv, wp), a.inv_n));
    sse_out = sse_t;
    if (WANT_GRAD) {
      const float drho_data = __fmul_rn(-__fsub_rn(__fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(0.5f, sse_t), e_rho), fvi), fvi),
                                                   __fmul_rn(0.5f, nb)), a.inv_bs);
      R[L.orho] = __fadd_rn(__fadd_rn(drho_data, __fmul_rn(__fdiv_rn(__fmul_rn(2.0f, dl), lv_den), a.inv_n)),
                            __fmul_rn(rho, pscale));
    }
  }
  if (!WANT_GRAD) {
    chain_barrier<NW>(bar_id);
    return;
  }
  // dW4: one column per thread (two for a single-warp chain)
  for (int col = tid; col <= HID; col += NTHR) {
    float sum = 0.0f;
#pragma unroll
    for (int ww = 0; ww < NW; ++ww) sum += scr[64 * ww + col];
    R[L.oW4 + col] = fmaf(R[L.oW4 + col], pscale, sum);
  }

  // ---- layers 3, 2 backward ----
#pragma unroll 1
  for (int l = 0; l < 2; ++l) {
    float* Wb = R + (l == 0 ? L.oW3 : L.oW2);
    const float* Hb = l == 0 ? s.Q : s.P;               // the activations below this layer
    store_c(s.Zb, batch, r0, t, h);                     // dZ of this layer, for the dW GEMM
    gemm_backward_data<MODE>(Wb, h, acc, g, t);
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) {
      const float2 v0 = *reinterpret_cast<const float2*>(Hb + min(r0, batch - 1) * AS + 8 * nt + 2 * t);
      const float2 v1 = *reinterpret_cast<const float2*>(Hb + min(r1, batch - 1) * AS + 8 * nt + 2 * t);
      const bool dead = nt == NT8 - 1 && t >= 1;        // columns >= 50
      h[nt][0] = (dead || r0 >= batch) ? 0.0f : acc[nt][0] * fmaf(-v0.x, v0.x, 1.0f);
      h[nt][1] = (dead || r0 >= batch) ? 0.0f : acc[nt][1] * fmaf(-v0.y, v0.y, 1.0f);
      h[nt][2] = (dead || r1 >= batch) ? 0.0f : acc[nt][2] * fmaf(-v1.x, v1.x, 1.0f);
      h[nt][3] = (dead || r1 >= batch) ? 0.0f : acc[nt][3] * fmaf(-v1.y, v1.y, 1.0f);
    }
    chain_barrier<NW>(bar_id);               // W of this layer is dead, Zb is complete
    gemm_weight_grad<NB8, 4 / NW, MODE>(Hb, s.Zb, Wb, batch, pscale, w * (4 / NW), g, t);
    chain_barrier<NW>(bar_id);               // Zb may be overwritten by the next dZ
  }

  // ---- layer 1 backward: dW1 = X^T dZ1, db1 = 1^T dZ1 (columns of dZ1 through Zb) ----
  store_c(s.Zb, batch, r0, t, h);
  chain_barrier<NW>(bar_id);
  for (int j = tid; j < HID; j += NTHR) {
    float db = 0.0f;
    for (int i = 0; i < batch; ++i) db += s.Zb[i * AS + j];
    R[L.ob1 + j] = fmaf(R[L.ob1 + j], pscale, db);
    for (int m = 0; m < n_in; ++m) {
      float dw = 0.0f;
      for (int i = 0; i < batch; ++i) dw = fmaf(s.sX[i * n_in + m], s.Zb[i * AS + j], dw);
      R[L.oW1 + m * HID + j] = fmaf(R[L.oW1 + m * HID + j], pscale, dw);
    }
  }
  chain_barrier<NW>(bar_id);
}

template <int NB8, bool WANT_GRAD, int MODE, int MINB = (NB8 > 2 ? 6 : 8), int BATCH_CT = 0>
__global__ void __launch_bounds__(32 * ((NB8 + 1) / 2), MINB) bnn_mma_kernel(BnnArgs a) {
  constexpr int NW = (NB8 + 1) / 2;
  constexpr int NTHR = 32 * NW;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int D = a.L.D;
  const BnnMmaSmem s = bnn_mma_carve(smem, a.batch, a.L.n_in, D);
  for (int64_t chain = blockIdx.x; chain < a.n_chains; chain += gridDim.x) {
    const float* th = a.theta + chain * D;
    float cost = 0.0f, sse = 0.0f;
    // Pull a LATER chain's parameter row towards L2 while this one computes: the row this CTA takes next
    // (persistent grid), or -- one CTA per chain -- the row of the chain about one wave of CTAs further on
    // (6 CTAs x 148 SMs are resident at a time; 1024 rows are 21 MB of the 126 MB L2).  Waiting for the row at
    // the head of a chain was the largest single stall of the kernel (16 % of the samples, ncu).
    {
      const int64_t ahead = gridDim.x < a.n_chains ? (int64_t)gridDim.x : (int64_t)K4_PREFETCH_AHEAD;
      if (tid == 0 && K4_PREFETCH_AHEAD > 0 && chain + ahead < a.n_chains && (D & 3) == 0 && aligned_to_dev(a.theta, 16))
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(th + ahead * D), "r"(D * 4) : "memory");
    }
    bnn_chain_mma<NB8, WANT_GRAD, false, MODE, BATCH_CT>(a, th, a.starts != nullptr ? a.starts + chain : nullptr, s, cost,
                                               sse);
    if (tid == 0) {
      a.cost[chain] = cost;
      if (a.mse != nullptr) a.mse[chain] = sse / (float)a.batch;
    }
    if (WANT_GRAD && a.grad != nullptr) {
      float* gr = a.grad + chain * D;
      if ((D & 3) == 0 && aligned_to_dev(gr, 16)) {
        const float4* src = reinterpret_cast<const float4*>(s.R);
        float4* dst = reinterpret_cast<float4*>(gr);
#pragma unroll 4
        for (int q = tid; q < D / 4; q += NTHR) dst[q] = src[q];
      } else {
        for (int q = tid; q < D; q += NTHR) gr[q] = s.R[q];
      }
    }
    chain_barrier<NW>();               // R is restaged by the next chain
  }
}

}  // namespace sgmcmc
