// Shared definitions of the BNN kernels (K4 bnn.cu / bnn_mma.cuh, K5 bnn_fused.cu, K10): the
// parameter layout of get_default_net (pysgmcmc/models/bayesian_neural_network.py:28-69) and
// the argument block of the cost + gradient kernels.
#pragma once

#include "sampler_math.cuh"

namespace sgmcmc {

constexpr int HID = 50;      // hidden width of get_default_net (bayesian_neural_network.py:30-49)
constexpr int HS = 52;       // row stride of the activation buffers: rows stay 16-byte aligned;
                             // columns 50, 51 are zero padding
constexpr int K4S = HS / 4;  // float4 steps per k-loop

struct BnnLayout {
  int n_in, D;
  int oW1, ob1, oW2, ob2, oW3, ob3, oW4, ob4, orho;
};

inline BnnLayout make_layout(int n_in) {
  BnnLayout L;
  L.n_in = n_in;
  int o = 0;
  L.oW1 = o; o += n_in * HID;
  L.ob1 = o; o += HID;
  L.oW2 = o; o += HID * HID;
  L.ob2 = o; o += HID;
  L.oW3 = o; o += HID * HID;
  L.ob3 = o; o += HID;
  L.oW4 = o; o += HID;
  L.ob4 = o; o += 1;
  L.orho = o; o += 1;
  L.D = o;
  return L;
}

struct BnnArgs {
  const float* theta;     // [C, D]
  const float* X;         // [N, n_in]
  const float* y;         // [N]
  const int32_t* starts;  // [C] (NULL: every chain starts at row 0)
  float* cost;            // [C]
  float* grad;            // [C, D] or NULL
  float* mse;             // [C] or NULL
  int64_t n_chains;
  int batch;              // rows actually in the minibatch
  float inv_bs;           // 1 / configured batch size          (:377)
  float inv_n;            // 1 / n_examples                     (:380)
  float prior_den_inv;    // 1 / (D + 3e-16)   safe_divide in weight_prior_log_like (:141)
  BnnLayout L;
};

// tanh(x) = 1 - 2 / (exp(2x) + 1) on the SFU (ex2.approx + rcp.approx): |error| ~ 2e-7,
// saturates correctly at +-1.  3000 activations per chain-step make the libm tanhf
// (~25 instructions) a third of the kernel; this is 5.  ex2.approx.ftz directly: __expf wraps the
// same instruction in a denormal-range fix-up (3 more instructions) that only matters for
// exp(2x) < 2^-126, where the result is -1 either way.
__device__ __forceinline__ float fast_tanh(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));   // 2 log2(e)
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}

// fast_tanh of two values with the FP32-pipe part packed (FMUL2 / FADD2 / FFMA2 around the four SFU
// instructions: 7 issue slots instead of 10); the same arithmetic, bit for bit
__device__ __forceinline__ void fast_tanh2(float x0, float x1, float& y0, float& y1) {
  asm("{\n\t.reg .b64 x, c, one, m2, r;\n\t.reg .f32 e0, e1;\n\t"
      "mov.b64 x, {%2, %3};\n\t"
      "mov.b64 c, {%4, %4};\n\t"
      "mov.b64 one, {%5, %5};\n\t"
      "mov.b64 m2, {%6, %6};\n\t"
      "mul.rn.f32x2 x, x, c;\n\t"
      "mov.b64 {e0, e1}, x;\n\t"
      "ex2.approx.ftz.f32 e0, e0;\n\t"
      "ex2.approx.ftz.f32 e1, e1;\n\t"
      "mov.b64 x, {e0, e1};\n\t"
      "add.rn.f32x2 x, x, one;\n\t"
      "mov.b64 {e0, e1}, x;\n\t"
      "rcp.approx.ftz.f32 e0, e0;\n\t"
      "rcp.approx.ftz.f32 e1, e1;\n\t"
      "mov.b64 r, {e0, e1};\n\t"
      "fma.rn.f32x2 r, r, m2, one;\n\t"
      "mov.b64 {%0, %1}, r;\n\t}"
      : "=f"(y0), "=f"(y1)
      : "f"(x0), "f"(x1), "f"(2.885390081777927f), "f"(1.0f), "f"(-2.0f));
}

__device__ __forceinline__ bool aligned_to_dev(const void* p, size_t a) {
  return (reinterpret_cast<uintptr_t>(p) % a) == 0;
}

// ---- K5 fused step (bnn_fused.cu) ----------------------------------------------------
struct FusedStepArgs {
  float *theta, *v, *tau, *g, *v_hat, *minv;   // [C, D] state (theta is read AND written)
  const float* z;                              // [C, D] injected noise of this step or NULL
  SghmcScalars<float> s;
  NoiseArgs na;
  int burn_in, store_minv;
  int prefetch;                                // TMA-prefetch the chain's state rows into L2
};
// true when the one-kernel step can run for this problem shape (else: K4 then K1)
bool bnn_fused_supported(const BnnArgs& a, const FusedStepArgs& f);
int launch_bnn_sghmc_fused(const BnnArgs& a, const FusedStepArgs& f, cudaStream_t st);
int bnn_fused_enabled();
void set_bnn_fused(int on);
void set_bnn_fused_max_ctas(int n);
void set_bnn_fused_prefetch(int on);

}  // namespace sgmcmc
