/*
 * sgmcmc_b200.h -- C ABI of libsgmcmc_b200.so, the B200 (sm_100a) SG-MCMC engine.
 *
 * This is the drop-in boundary for the sampler hot path of MFreidank/pysgmcmc.
 * The reference has no FFI of its own (pure Python on TensorFlow 1.x); each
 * entry point below replaces the TensorFlow graph region cited next to it and
 * is what a binding for that path would call (ctypes stub: INTEGRATION.md).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller; the library never
 *     allocates or frees caller-visible memory and keeps no global state besides
 *     a thread-local error string and the launch-tuning knobs;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     all calls are asynchronous and stream ordered;
 *   - state is laid out [C chains x D params] row-major, flattened; `n` = C*D;
 *   - return value: 0 on success, <0 on error (SGMCMC_E_*), message via
 *     sgmcmc_last_error();
 *   - noise: `z` != NULL -> N(0,1) draws are READ from z (same layout as theta);
 *     `z` == NULL -> generated in-kernel: Philox4x32-10, counter =
 *     ((elem_offset+e)/4, step), key = seed, Box-Muller (see oracle/philox.py).
 *     `elem_offset` (multiple of 4) is the global flat index of element 0, so a
 *     shard of chains reproduces the stream of the un-sharded run.
 */
#ifndef SGMCMC_B200_H
#define SGMCMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGMCMC_OK 0
#define SGMCMC_E_INVALID (-1)   /* bad argument (NULL pointer, negative size, ...) */
#define SGMCMC_E_ALIGN (-2)     /* pointer not aligned to the element size       */
#define SGMCMC_E_CUDA (-3)      /* CUDA launch / runtime error                   */
#define SGMCMC_E_UNSUPPORTED (-4)

#define SGMCMC_SAMPLER_SGHMC 0
#define SGMCMC_SAMPLER_SGLD 1
#define SGMCMC_SAMPLER_RSGHMC 2

#define SGMCMC_TARGET_BANANA 0  /* diagnostics/objective_functions.py:49-59 */
#define SGMCMC_TARGET_GMM1 1    /* :89-90 */
#define SGMCMC_TARGET_GMM2 2    /* :93-94 */
#define SGMCMC_TARGET_GMM3 3    /* :97-98 */

int sgmcmc_version(void);
const char* sgmcmc_last_error(void);

/* Launch tuning for the element-wise update kernels (threads per CTA in
 * {128,256,512}, float4 groups per thread in {1,2}); 0 keeps the current value.
 * Measured on B200 (profiles/): 256 x 1 is fastest; more groups per thread cost occupancy. */
int sgmcmc_set_update_tuning(int threads, int unroll);

/* Implementation of the BNN kernel K4: 10-16 are the tensor-pipe kernel (3xTF32 mma.sync,
 * csrc/bnn_mma.cuh; minibatches of up to 32 rows, larger ones fall back to 0) in its accuracy
 * modes (10: truncating hi/lo split, sums chained through the tensor core's accumulator;
 * 11: rounded split; 12: FP32-pipe accumulation across k-steps; 13: both; 14: 13 with the weight
 * fragments split by packed FP32 instructions; 15: 13 with the cross terms of the 3xTF32 products in their
 * own accumulator; 16: 15 + 14, the default); 0-9 are launch
 * shapes of the FFMA kernel (units per thread, chains per CTA, rows in flight) kept for the
 * sweeps recorded under profiles/. */
int sgmcmc_set_bnn_tuning(int variant);

/* K1 walks its arrays from the end (the first CTAs take the last elements): after K4, which
 * walks the chains in ascending order, the tail of theta and of the gradient is still in L2
 * (and K1 then ends where the next K4 begins).  On by default; results do not depend on it. */
int sgmcmc_set_update_reverse(int on);

/* Cap the grids of the update kernels (K1-K3) and of K4 to that many CTAs (persistent
 * kernels that loop over their work; 0 = one CTA per unit of work, the default).  Used to
 * leave SM resources free when two kernels are meant to run concurrently on two streams. */
int sgmcmc_set_persistent_grids(int update_max_ctas, int bnn_max_ctas);

/* Chains per chunk inside sgmcmc_bnn_sghmc_run_f32: K4 and K1 run back to back on one chunk
 * at a time so that the chunk's gradient stays in L2 (0 = all chains in one chunk). */
int sgmcmc_set_bnn_chunk(int64_t chains);

/* Two-stream pipeline inside sgmcmc_bnn_sghmc_run_f32: the chains are walked in chunks of
 * `chunk_chains`; K4 of a chunk runs on the caller's stream while K1 of the previous chunk
 * runs on a stream owned by the library, the gradient going through a ring of `ring` (default
 * 2) chunk-sized slots of grad_scratch.  Results are bit-identical to the sequential order.
 * chunk_chains = 0 switches it off. */
int sgmcmc_set_bnn_pipeline(int64_t chunk_chains, int ring);

/* sgmcmc_bnn_sghmc_run_f32 as ONE kernel per step (cost + gradient + SGHMC update of a chain
 * in one CTA, the gradient never leaves shared memory; csrc/bnn_fused.cu): bit-identical to
 * K4 then K1, 40 instead of 52 B of HBM traffic per element-step and no grad_scratch, but
 * slower at large chain counts (too few warps per SM for the update's instruction stream,
 * DESIGN.md "K5"), so on = 0 (K4 then K1) is the default.  on = 1 selects the fused kernel,
 * on = 3 the fused kernel without the TMA L2 prefetch of the state rows, on = 4 (6: without the
 * prefetch) its warp-specialised form: persistent CTAs of two 2-warp MMA groups and four update warps,
 * the gradient handed over through two shared-memory buffers per group, so the issue-bound gradient and the
 * HBM-bound update overlap inside every SM.  max_ctas > 0 caps the grid (persistent CTAs looping over
 * chains; the warp-specialised kernel is always persistent, 2 CTAs per SM by default). */
int sgmcmc_set_bnn_fused(int on, int max_ctas);

/* Number of kernel launches issued by this library since load (all threads). */
int64_t sgmcmc_launch_count(void);

/* ---- K1: SGHMC update, replaces pysgmcmc/samplers/sghmc.py:165-251 ----------------
 * burn_in != 0: adapts tau/g/v_hat and computes minv from the OLD v_hat; if
 *               store_minv != 0 the minv used is also written to `minv`.
 * burn_in == 0: `minv` is READ (the mass matrix frozen by
 *               samplers/base_classes.py:448-454); tau/g/v_hat are not touched.
 * grad = d cost / d theta at the old theta.  epsilon is the UNSCALED step size. */
int sgmcmc_sghmc_step_f32(float* theta, float* v, float* tau, float* g, float* v_hat, float* minv,
                          const float* grad, const float* z, int64_t n,
                          float epsilon, float mdecay, float scale_grad,
                          int burn_in, int store_minv,
                          uint64_t seed, uint64_t step, uint64_t elem_offset, void* stream);
int sgmcmc_sghmc_step_f64(double* theta, double* v, double* tau, double* g, double* v_hat, double* minv,
                          const double* grad, const double* z, int64_t n,
                          double epsilon, double mdecay, double scale_grad,
                          int burn_in, int store_minv,
                          uint64_t seed, uint64_t step, uint64_t elem_offset, void* stream);

/* ---- K2: SGLD update, replaces pysgmcmc/samplers/sgld.py:149-213 ------------------ */
int sgmcmc_sgld_step_f32(float* theta, float* tau, float* g, float* v_hat, float* minv,
                         const float* grad, const float* z, int64_t n,
                         float epsilon, float A, float scale_grad,
                         int burn_in, int store_minv,
                         uint64_t seed, uint64_t step, uint64_t elem_offset, void* stream);
int sgmcmc_sgld_step_f64(double* theta, double* tau, double* g, double* v_hat, double* minv,
                         const double* grad, const double* z, int64_t n,
                         double epsilon, double A, double scale_grad,
                         int burn_in, int store_minv,
                         uint64_t seed, uint64_t step, uint64_t elem_offset, void* stream);

/* ---- K3: relativistic SGHMC update, replaces
 * pysgmcmc/samplers/relativistic_sghmc.py:120-140 (element-wise momentum).
 * grad_cost = d cost / d theta (the kernel negates it, :100-103). */
int sgmcmc_rsghmc_step_f32(float* theta, float* p, const float* grad_cost, const float* z, int64_t n,
                           float epsilon, float mass, float speed_of_light, float D, float Bhat,
                           uint64_t seed, uint64_t step, uint64_t elem_offset, void* stream);
int sgmcmc_rsghmc_step_f64(double* theta, double* p, const double* grad_cost, const double* z, int64_t n,
                           double epsilon, double mass, double speed_of_light, double D, double Bhat,
                           uint64_t seed, uint64_t step, uint64_t elem_offset, void* stream);

/* The engine's N(0,1) stream written out (replaces tf.random_normal,
 * samplers/base_classes.py:218-220; also the test hook for the in-kernel noise). */
int sgmcmc_normal_fill_f32(float* out, int64_t n, uint64_t seed, uint64_t step,
                           uint64_t elem_offset, void* stream);

/* ---- K6: whole chains on a built-in target, many steps per launch -----------------
 * One thread owns one chain (D = 2 for banana, 1 for gmm*): gradient of the target
 * (diagnostics/objective_functions.py:49-98), the sampler update and the noise are
 * fused; `n_steps` steps run inside one launch with the state in registers.
 * Steps [0, n_burn_in) adapt (and the last of them stores minv), the rest use the
 * frozen minv; adapt_forever != 0 reproduces burn_in_steps == 0
 * (samplers/base_classes.py:449).  Unused state pointers may be NULL
 * (SGLD: a1; RSGHMC: tau,g,v_hat,minv; a1 is V for SGHMC and p for RSGHMC).
 * z: NULL or [n_steps, C, D].  Outputs (either may be NULL): every keep_every-th
 * step s (s+1 divisible by keep_every) writes theta (post-update) to
 * trace[(s+1)/keep_every-1, C, D] and the cost at the pre-update point to
 * cost_trace[(s+1)/keep_every-1, C]  -- the (sample, cost) pair next(sampler)
 * returns (samplers/base_classes.py:298-300). */
typedef struct {
  float epsilon;
  float mdecay;          /* SGHMC */
  float scale_grad;      /* SGHMC, SGLD */
  float A;               /* SGLD */
  float mass, speed_of_light, D, Bhat;   /* RSGHMC */
} sgmcmc_hyper_t;

int sgmcmc_target_chains_run_f32(int sampler, int target,
                                 float* theta, float* a1, float* tau, float* g, float* v_hat,
                                 float* minv, const float* z, float* trace, float* cost_trace,
                                 int64_t n_chains, int64_t n_steps, int64_t n_burn_in,
                                 int adapt_forever, int64_t keep_every,
                                 const sgmcmc_hyper_t* hyper,
                                 uint64_t seed, uint64_t step0, uint64_t chain_offset, void* stream);

/* ---- K7: minibatch start indices, replaces pysgmcmc/data_batches.py:104-120 -------
 * One MT19937 stream per chain, bit-exact with numpy.random.RandomState(seed).
 * state: uint32 [625, n_streams] (624 words + position, stream-minor). */
int sgmcmc_mt19937_seed(uint32_t* state, const uint32_t* seeds, int64_t n_streams, void* stream);
/* starts[s, j] = the (s+1)-th future value of rng_j.randint(0, max_inclusive + 1). */
int sgmcmc_mt19937_starts(uint32_t* state, int32_t* starts, int64_t n_streams, int64_t n_steps,
                          uint32_t max_inclusive, void* stream);

/* ---- K4: BNN cost and gradient, replaces
 * pysgmcmc/models/bayesian_neural_network.py:28-69 (get_default_net), :77-141
 * (priors), :337-388 (negative_log_likelihood) and tf.gradients over it.
 * Network n_in -> 50 -> 50 -> 50 -> 1 (tanh), flat per-chain layout
 * W1 b1 W2 b2 W3 b3 W4 b4 rho (D = 50*n_in + 5202).  Chain j reads the minibatch
 * rows X[starts[j] : starts[j]+batch], y[...] (data_batches.py:120-123).
 * batch_size_cfg is the constant the reference divides the data term by (:377).
 * cost [C]; grad [C, D] (may be NULL); mse [C] (may be NULL). */
int sgmcmc_bnn_nll_grad_f32(const float* theta, const float* X, const float* y,
                            const int32_t* starts, float* cost, float* grad, float* mse,
                            int64_t n_chains, int n_in, int batch, float batch_size_cfg,
                            int64_t n_examples, void* stream);

/* ---- K4 / K10 for any fully connected `get_net`: n_in -> h_1 -> ... -> h_L -> 1 with tanh
 * hidden layers, a linear head and the learned log variance -- the architecture family of
 * get_default_net (bayesian_neural_network.py:28-69) with user-chosen widths and depth, e.g. the
 * 1000-512-512 network of BASELINE.json configs[4].  widths = {n_in, h_1, ..., h_L, 1}
 * (n_widths = L + 2, 1 <= L <= 7).  Flat per-chain layout W_1 b_1 ... W_{L+1} b_{L+1} rho
 * (sgmcmc_mlp_n_params values; kernels [in, out] row-major).  Cost, priors and gradient as
 * sgmcmc_bnn_nll_grad_f32 (:77-141, :337-388); minibatches of up to 32 rows.  `workspace`
 * (device, 16-byte aligned, sgmcmc_mlp_workspace_bytes(widths, n_widths, n_chains, batch) bytes)
 * holds the activations between the layer kernels (csrc/mlp.cu).
 * sgmcmc_mlp_predict_f32: out[k, i, 0] = f(x_i; theta_k), out[k, i, 1] = rho_k (:535-557);
 * its workspace is sized with n_items = n_nets * ceil(n_points / 32), batch = 32. */
/* sgmcmc_set_mlp_tuning(1) (the default): layers whose weight matrix is at least 128 x 128 (widths
 * multiples of 4) run their forward and backward-data GEMMs on the tcgen05 tensor cores as 3xTF32
 * (csrc/mlp_umma.cu); 0 keeps every layer on the FFMA kernels (the implementation the tests compare
 * against).  Set it BEFORE sizing the workspace: the tensor-core layers keep split copies there. */
int sgmcmc_set_mlp_tuning(int tensor_core_layers);
int64_t sgmcmc_mlp_n_params(const int* widths, int n_widths);
int64_t sgmcmc_mlp_workspace_bytes(const int* widths, int n_widths, int64_t n_items, int batch);
int sgmcmc_mlp_nll_grad_f32(const float* theta, const float* X, const float* y, const int32_t* starts,
                            float* cost, float* grad, float* mse, void* workspace, int64_t workspace_bytes,
                            int64_t n_chains, const int* widths, int n_widths, int batch,
                            float batch_size_cfg, int64_t n_examples, void* stream);
int sgmcmc_mlp_predict_f32(const float* theta, const float* X, float* out, void* workspace,
                           int64_t workspace_bytes, int64_t n_nets, const int* widths, int n_widths,
                           int64_t n_points, void* stream);

/* ---- K5: BNN-SGHMC chains: K4 + K1 for `n_steps` steps in one call with no host
 * synchronisation -- the whole next(sampler) of the BNN path
 * (samplers/base_classes.py:408-456 driving sghmc.py:165-251 over
 * bayesian_neural_network.py:337-388).
 * starts: int32 [n_steps, C] (from sgmcmc_mt19937_starts; NULL: every minibatch starts at
 * row 0).  Same burn-in / noise / trace conventions as sgmcmc_target_chains_run_f32; z is
 * NULL or [n_steps, C, D].  grad_scratch [C, D] and cost_scratch [C] are caller-owned
 * work buffers (the gradient and cost of the last step are left in them). */
int sgmcmc_bnn_sghmc_run_f32(float* theta, float* v, float* tau, float* g, float* v_hat, float* minv,
                             const float* X, const float* y, const int32_t* starts,
                             const float* z, float* trace, float* cost_trace,
                             float* grad_scratch, float* cost_scratch,
                             int64_t n_chains, int n_in, int batch, float batch_size_cfg,
                             int64_t n_examples, int64_t n_steps, int64_t n_burn_in,
                             int adapt_forever, int64_t keep_every,
                             float epsilon, float mdecay, float scale_grad,
                             uint64_t seed, uint64_t step0, uint64_t chain_offset, void* stream);

/* ---- K5r: the same `n_steps` steps with every chain RESIDENT on one SM (csrc/bnn_resident.cu): one CTA per
 * chain loads theta, V, tau, g, v_hat, minv into shared memory once, runs all n_steps steps there (cost +
 * gradient on the FP32 pipe in the accumulation order of K4's FFMA kernel, then K1's arithmetic with K1's
 * Philox counters) and writes the state back: HBM traffic per chain and CALL instead of per chain and step,
 * and no launch per step -- the path for few chains (the reference's single-chain BOHAMIANN runs,
 * bayesian_neural_network.py:436-531); faster than K4 then K1 up to a few chains per SM, slower from ~1000
 * chains on (DESIGN.md "K5r").  Chains never interact, so the
 * result does not depend on how a run is cut into calls.  Needs sgmcmc_bnn_resident_supported(n_in, batch)
 * (batch <= 32, state + activations within 227 KB); E_UNSUPPORTED otherwise.  The noise of element group q of
 * chain c comes from Philox counter (chain_offset + c) * ceil(D / 4) + q: K1's counters when D % 4 == 0 (odd n_in).
 * cost_last [C] receives the cost of the last step; cost_all (NULL or [n_steps, C]) the cost of every step;
 * grad_out (NULL or [C, D]) the gradient of the last step; the other arguments as in
 * sgmcmc_bnn_sghmc_run_f32.  `minv` is defined from the last burn-in step on (frozen there); during burn-in it
 * may already hold the inverse mass matrix of the latest step (it does with the overlap below).  sgmcmc_set_bnn_resident_overlap(0): the whole
 * update runs after the gradient instead of its gradient-free part beside it (what happens anyway when the two
 * extra D-float arrays do not fit shared memory: minibatch > 20 rows); same bits, for the measurements in profiles/. */
int sgmcmc_bnn_resident_supported(int n_in, int batch);
int sgmcmc_set_bnn_resident_overlap(int on);
int sgmcmc_bnn_sghmc_run_resident_f32(float* theta, float* v, float* tau, float* g, float* v_hat, float* minv,
                                      const float* X, const float* y, const int32_t* starts,
                                      const float* z, float* trace, float* cost_trace,
                                      float* cost_all, float* cost_last, float* grad_out,
                                      int64_t n_chains, int n_in, int batch, float batch_size_cfg,
                                      int64_t n_examples, int64_t n_steps, int64_t n_burn_in,
                                      int adapt_forever, int64_t keep_every,
                                      float epsilon, float mdecay, float scale_grad,
                                      uint64_t seed, uint64_t step0, uint64_t chain_offset, void* stream);

/* ---- K5 with HOST buffers: the pipelined `sample, cost = next(sampler)` of the BNN path
 * (samplers/base_classes.py:258-310,408-456; csrc/host_pipeline.cu).  Every `step` copies
 * that step's minibatch start indices from pinned host memory (host_starts [C]), runs
 * K4 + K1 on `stream`, copies the per-chain cost to pinned host memory (host_cost [C]) and,
 * if host_sample != NULL, the new sample theta [C, D] too (through a device-side snapshot,
 * so later steps do not wait for PCIe).  `step` never blocks the host; `wait(ticket)` blocks
 * until the results of that step are in host memory.  Up to `depth` steps may be in flight:
 * a ticket can be waited for until `depth` further steps were enqueued, and its host buffers
 * must stay untouched until then.  burn_in_left = burn-in steps left including this one
 * (0 = sampling phase; minv is written back only on the last burn-in step).  The handle owns
 * the copy streams, events and device-side slot buffers.  with_samples: bit 0 = allocate the sample
 * stage, bit 1 = every step is one launch of the resident kernel (sgmcmc_bnn_sghmc_run_resident_f32,
 * grad_scratch unused) instead of K4 then K1 -- the arithmetic a sampler with few chains uses everywhere. */
typedef struct sgmcmc_bnn_host_pipeline sgmcmc_bnn_host_pipeline;
int sgmcmc_bnn_host_pipeline_create(sgmcmc_bnn_host_pipeline** out, int64_t n_chains, int n_in,
                                    int depth, int with_samples);
int sgmcmc_bnn_host_pipeline_destroy(sgmcmc_bnn_host_pipeline* p);
int sgmcmc_bnn_host_pipeline_step(sgmcmc_bnn_host_pipeline* p,
                                  float* theta, float* v, float* tau, float* g, float* v_hat, float* minv,
                                  const float* X, const float* y,
                                  const int32_t* host_starts, float* host_cost, float* host_sample,
                                  float* grad_scratch, int n_in, int batch, float batch_size_cfg,
                                  int64_t n_examples, int burn_in_left, int adapt_forever,
                                  float epsilon, float mdecay, float scale_grad,
                                  uint64_t seed, uint64_t step, uint64_t chain_offset, void* stream,
                                  int64_t* ticket);
int sgmcmc_bnn_host_pipeline_wait(sgmcmc_bnn_host_pipeline* p, int64_t ticket);

/* ---- K10: BNN predictive, replaces bayesian_neural_network.py:535-557 -------------
 * out[k, i, 0] = f(x_i; theta_k), out[k, i, 1] = rho_k  for n_nets stored samples. */
int sgmcmc_bnn_predict_f32(const float* theta, const float* X, float* out,
                           int64_t n_nets, int n_in, int64_t n_points, void* stream);

/* ---- K8: per-chain moments and lagged variogram sums for R-hat / ESS --------------
 * (formulas: pysgmcmc/diagnostics/sampler_diagnostics.py:76-82,153-161).
 * trace: float [n_draws, C, D].  sums: double [3, D] = sum_j mean_j, sum_j mean_j^2,
 * sum_j var_j (ddof=1) over the C local chains -- the quantities all-reduced across
 * GPUs.  variogram: double [n_lags, D], sum over chains and draws of
 * (x[i] - x[i-t])^2 for t = lag0 .. lag0+n_lags-1.  Either output may be NULL. */
int sgmcmc_chain_moments_f32(const float* trace, double* sums, int64_t n_draws, int64_t n_chains,
                             int64_t n_dims, void* stream);
int sgmcmc_variogram_f32(const float* trace, double* variogram, int64_t n_draws, int64_t n_chains,
                         int64_t n_dims, int64_t lag0, int64_t n_lags, void* stream);
/* The same sums for the n_sel dimensions listed in `dims` (device int64 [n_sel]) only:
 * variogram is double [n_lags, n_sel].  Used for the few dimensions whose ESS stopping rule has
 * not fired after the first block of lags. */
int sgmcmc_variogram_select_f32(const float* trace, const int64_t* dims, double* variogram,
                                int64_t n_draws, int64_t n_chains, int64_t n_dims, int64_t n_sel,
                                int64_t lag0, int64_t n_lags, void* stream);

/* ---- Stein variational gradient descent (K11-K14) -------------------------------------
 * Replaces pysgmcmc/samplers/svgd.py:81-182 and the helpers it calls,
 * pysgmcmc/tensor_utils.py:160-208 (median), :326-419 (pdist), :422-577 (squareform).
 * The particles are the chain layout: float [n_particles, n_dims] row-major.
 *
 * sgmcmc_median_f32: out[0] = median of `values` (middle value, or the mean of the two
 * middle values of an even count; tensor_utils.py:194-208).  Exact (radix select), any
 * finite floats.  scratch: 4096 bytes of device memory, 8-byte aligned.
 *
 * sgmcmc_svgd_kernel_matrix_f32 (svgd.py:150-160): kernel_matrix [n, n] =
 * exp(-P / h^2 / 2) with P[i,j] = (||x_i - x_j||)^2 and h = sqrt(0.5 * median(P) /
 * log(n + 1)); kernel_sum [n] = its row sums; bandwidth [4] = {median(P), h, h^2, 0}.
 * Everything stays on the device.  n_particles <= 46340.  scratch: device memory, 16-byte
 * aligned, `scratch_bytes` long: 4096 bytes always; 4096 + 4 * (n_particles + n_dims + 3) bytes
 * to be eligible for the tensor-core distance kernel; sgmcmc_svgd_scratch_bytes() returns the size
 * with which that kernel may also slice the contraction over the dimensions when the particle
 * count alone gives too few tiles to fill the SMs.  kernel_matrix comes out symmetric bit for
 * bit (sgmcmc_svgd_update_f32 reads its rows as columns).  Large aligned shapes compute the
 * distances from the Gram matrix of the centred particles on the tcgen05 tensor cores
 * (3xTF32, csrc/svgd_sqdist_umma.cu), the others subtract before squaring on the FP32 pipe.
 *
 * sgmcmc_svgd_update_f32 (svgd.py:125-148,162-167): with grad[i] = d cost(x_i) / d x_i,
 *   phi   = (K @ grad + (-(K @ X) + X * kernel_sum[:, None]) / h^2) / n
 *   hist  = alpha * hist + one_minus_alpha * phi^2
 *   X    -= epsilon * phi / (fudge_factor + sqrt(hist))
 * historical_grad and particles are updated in place (particles through
 * particles_scratch [n, D], because every output row reads all of X). */
/* Implementation of the distance kernel (K11) and of the update GEMM (K14): 0 = automatic
 * (default: tensor cores for n_particles >= 256 / 128 and n_dims >= 128), 1 = FP32 FFMA
 * kernels, 2 = tcgen05 tensor-core kernels (3xTF32 products accumulated in TMEM,
 * csrc/svgd_umma.cu, csrc/svgd_sqdist_umma.cu) whenever the shape is eligible (n_dims % 4 == 0,
 * for K14 also n_particles % 4 == 0, 16-byte aligned pointers), otherwise the FFMA kernels run;
 * 23 / 24 = 2 with 3 / 4 producer register buffers in K14 (sweeps). */
int sgmcmc_set_svgd_tuning(int impl);
int sgmcmc_median_f32(const float* values, int64_t n_values, float* out, void* scratch, void* stream);
/* The same median over all n * n entries of a SYMMETRIC matrix with a ZERO diagonal (what the
 * kernel-matrix entry point uses on the squared distances): only the upper triangle is read. */
int sgmcmc_median_symmetric_f32(const float* matrix, int64_t n, float* out, void* scratch, void* stream);
int64_t sgmcmc_svgd_scratch_bytes(int64_t n_particles, int64_t n_dims);
int sgmcmc_svgd_kernel_matrix_f32(const float* particles, float* kernel_matrix, float* kernel_sum,
                                  float* bandwidth, void* scratch, int64_t scratch_bytes,
                                  int64_t n_particles, int64_t n_dims, void* stream);
/* SVGD of at most 128 particles on a built-in test density (SGMCMC_TARGET_*), `n_steps` steps in
 * ONE launch of one CTA (gradients, distances, exact median, kernel, Stein direction and update
 * all in shared memory): the regime of docs/source/notebooks/SVGD.ipynb.  particles
 * [n, D] (D = 2 banana, 1 gmm) and historical_grad [n, D] are updated in place; every
 * keep_every-th particle set goes to trace [n_steps / keep_every, n, D] and the costs of the
 * particles BEFORE that step to cost_trace [n_steps / keep_every, n] (either may be NULL). */
int sgmcmc_svgd_target_run_f32(int target, float* particles, float* historical_grad, float* trace,
                               float* cost_trace, int64_t n_particles, int64_t n_steps,
                               int64_t keep_every, float epsilon, float alpha, float one_minus_alpha,
                               float fudge_factor, void* stream);
int sgmcmc_svgd_update_f32(float* particles, const float* grad, float* historical_grad,
                           const float* kernel_matrix, const float* kernel_sum, const float* bandwidth,
                           float* particles_scratch, int64_t n_particles, int64_t n_dims, float epsilon,
                           float alpha, float one_minus_alpha, float fudge_factor, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SGMCMC_B200_H */
