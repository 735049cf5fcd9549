// The host-facing stepper of the BNN-SGHMC path: `sample, cost = next(sampler)`
// (pysgmcmc/samplers/base_classes.py:258-310,408-456) with the minibatch choice coming from
// HOST memory every step and the cost (and every n-th sample) going back to HOST memory, as a
// pipeline that never lets the device wait for the host:
//
//   stream `in`  : H2D of the step's start indices into slot b = ticket % depth
//   caller stream: K4 + K1 of the step (sgmcmc_bnn_sghmc_run_f32), a device-to-device
//                  snapshot of theta when the sample is wanted
//   stream `out` : D2H of the per-chain cost (and of the snapshot)
//
// Slots are recycled with events only (no host synchronisation inside `step`): the H2D into
// slot b waits for the kernels that last read it, the kernels wait for the D2H that last
// read the slot's cost buffer.  The host blocks in `wait(ticket)` only, for results it is
// about to hand to the caller, while up to depth - 1 younger steps are already queued.
// The device-side slot buffers are owned by the handle (create / destroy pair).
#include <new>

#include "bnn_common.cuh"

namespace sgmcmc {

constexpr int MAX_DEPTH = 64;

}  // namespace sgmcmc

struct sgmcmc_bnn_host_pipeline {
  int64_t n_chains = 0, D = 0;
  int depth = 0;
  int resident = 0;               // steps through the resident kernel (bnn_resident.cu) instead of K4 then K1
  int64_t next_ticket = 0;
  int64_t n_samples = 0;          // samples requested so far (stage buffer = n_samples % 2)
  cudaStream_t s_in = nullptr, s_out = nullptr;
  int32_t* d_starts = nullptr;   // [depth, C]
  float* d_cost = nullptr;       // [depth, C]
  float* d_stage = nullptr;      // [2, C, D] or NULL: two snapshots may be on their way out
  cudaEvent_t h2d[sgmcmc::MAX_DEPTH], step_done[sgmcmc::MAX_DEPTH], out_done[sgmcmc::MAX_DEPTH];
  cudaEvent_t sample_done[2] = {nullptr, nullptr};
};

using namespace sgmcmc;

#define SG_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) return set_error(SGMCMC_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

extern "C" int sgmcmc_bnn_host_pipeline_destroy(sgmcmc_bnn_host_pipeline* p) {
  if (p == nullptr) return SGMCMC_OK;
  if (p->s_out) cudaStreamSynchronize(p->s_out);
  if (p->s_in) cudaStreamSynchronize(p->s_in);
  for (int b = 0; b < p->depth; ++b) {
    if (p->h2d[b]) cudaEventDestroy(p->h2d[b]);
    if (p->step_done[b]) cudaEventDestroy(p->step_done[b]);
    if (p->out_done[b]) cudaEventDestroy(p->out_done[b]);
  }
  for (int k = 0; k < 2; ++k)
    if (p->sample_done[k]) cudaEventDestroy(p->sample_done[k]);
  if (p->d_starts) cudaFree(p->d_starts);
  if (p->d_cost) cudaFree(p->d_cost);
  if (p->d_stage) cudaFree(p->d_stage);
  if (p->s_in) cudaStreamDestroy(p->s_in);
  if (p->s_out) cudaStreamDestroy(p->s_out);
  delete p;
  return SGMCMC_OK;
}

extern "C" int sgmcmc_bnn_host_pipeline_create(sgmcmc_bnn_host_pipeline** out, int64_t n_chains, int n_in,
                                               int depth, int with_samples) {
  SG_REQUIRE(out != nullptr, SGMCMC_E_INVALID, "host_pipeline_create: out must not be NULL");
  *out = nullptr;
  SG_REQUIRE(n_chains >= 1, SGMCMC_E_INVALID, "host_pipeline_create: n_chains must be >= 1");
  SG_REQUIRE(n_in >= 1 && n_in <= 64, SGMCMC_E_UNSUPPORTED, "bnn: n_in must be in [1, 64] (got %d)", n_in);
  SG_REQUIRE(depth >= 1 && depth <= MAX_DEPTH, SGMCMC_E_INVALID, "host_pipeline_create: depth must be in [1, %d]",
             MAX_DEPTH);
  sgmcmc_bnn_host_pipeline* p = new (std::nothrow) sgmcmc_bnn_host_pipeline();
  SG_REQUIRE(p != nullptr, SGMCMC_E_INVALID, "host_pipeline_create: out of host memory");
  for (int b = 0; b < MAX_DEPTH; ++b) p->h2d[b] = p->step_done[b] = p->out_done[b] = nullptr;
  p->n_chains = n_chains;
  p->D = make_layout(n_in).D;
  p->depth = depth;
  p->resident = (with_samples & 2) != 0;
  int rc = SGMCMC_OK;
  auto ok = [&](cudaError_t e, const char* what) {
    if (e != cudaSuccess && rc == SGMCMC_OK) rc = set_error(SGMCMC_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return e == cudaSuccess;
  };
  ok(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking), "cudaStreamCreate");
  ok(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking), "cudaStreamCreate");
  ok(cudaMalloc(&p->d_starts, sizeof(int32_t) * depth * n_chains), "cudaMalloc(starts)");
  ok(cudaMalloc(&p->d_cost, sizeof(float) * depth * n_chains), "cudaMalloc(cost)");
  if (with_samples & 1) ok(cudaMalloc(&p->d_stage, sizeof(float) * 2 * n_chains * p->D), "cudaMalloc(sample stage)");
  for (int b = 0; b < depth; ++b) {
    ok(cudaEventCreateWithFlags(&p->h2d[b], cudaEventDisableTiming), "cudaEventCreate");
    ok(cudaEventCreateWithFlags(&p->step_done[b], cudaEventDisableTiming), "cudaEventCreate");
    ok(cudaEventCreateWithFlags(&p->out_done[b], cudaEventDisableTiming), "cudaEventCreate");
  }
  for (int k = 0; k < 2; ++k)
    ok(cudaEventCreateWithFlags(&p->sample_done[k], cudaEventDisableTiming), "cudaEventCreate");
  if (rc != SGMCMC_OK) {
    sgmcmc_bnn_host_pipeline_destroy(p);
    return rc;
  }
  *out = p;
  return SGMCMC_OK;
}

extern "C" int sgmcmc_bnn_host_pipeline_step(sgmcmc_bnn_host_pipeline* p, float* theta, float* v, float* tau,
                                             float* g, float* v_hat, float* minv, const float* X,
                                             const float* y, const int32_t* host_starts, float* host_cost,
                                             float* host_sample, float* grad_scratch, int n_in, int batch,
                                             float batch_size_cfg, int64_t n_examples, int burn_in_left,
                                             int adapt_forever, float epsilon, float mdecay,
                                             float scale_grad, uint64_t seed, uint64_t step,
                                             uint64_t chain_offset, void* stream, int64_t* ticket) {
  SG_REQUIRE(p != nullptr && host_starts != nullptr && host_cost != nullptr, SGMCMC_E_INVALID,
             "host_pipeline_step: handle, host_starts and host_cost must not be NULL");
  SG_REQUIRE(host_sample == nullptr || p->d_stage != nullptr, SGMCMC_E_INVALID,
             "host_pipeline_step: the pipeline was created without a sample stage");
  SG_REQUIRE(make_layout(n_in).D == p->D, SGMCMC_E_INVALID, "host_pipeline_step: n_in differs from create()");
  SG_REQUIRE(burn_in_left >= 0, SGMCMC_E_INVALID, "host_pipeline_step: burn_in_left must be >= 0");
  cudaStream_t main = (cudaStream_t)stream;
  const int64_t C = p->n_chains;
  const int b = (int)(p->next_ticket % p->depth);
  const bool recycled = p->next_ticket >= p->depth;
  int32_t* d_starts = p->d_starts + (int64_t)b * C;
  float* d_cost = p->d_cost + (int64_t)b * C;
  if (recycled) SG_CUDA(cudaStreamWaitEvent(p->s_in, p->step_done[b], 0));     // slot's indices no longer read
  SG_CUDA(cudaMemcpyAsync(d_starts, host_starts, sizeof(int32_t) * C, cudaMemcpyHostToDevice, p->s_in));
  SG_CUDA(cudaEventRecord(p->h2d[b], p->s_in));
  SG_CUDA(cudaStreamWaitEvent(main, p->h2d[b], 0));
  if (recycled) SG_CUDA(cudaStreamWaitEvent(main, p->out_done[b], 0));         // slot's cost has left the device
  // n_burn_in of the one-step run: 0 (sampling), 1 (the LAST burn-in step: minv is written
  // back and frozen) or 2 (burn-in goes on: no write-back)
  const int64_t n_burn_in = burn_in_left > 2 ? 2 : burn_in_left;
  if (p->resident) {
    if (int rc = sgmcmc_bnn_sghmc_run_resident_f32(theta, v, tau, g, v_hat, minv, X, y, d_starts, nullptr, nullptr,
                                                   nullptr, nullptr, d_cost, nullptr, C, n_in, batch, batch_size_cfg,
                                                   n_examples, 1, burn_in_left > 0 ? 1 : 0, adapt_forever, 1, epsilon,
                                                   mdecay, scale_grad, seed, step, chain_offset, stream))
      return rc;
  } else if (int rc = sgmcmc_bnn_sghmc_run_f32(theta, v, tau, g, v_hat, minv, X, y, d_starts, nullptr, nullptr, nullptr,
                                        grad_scratch, d_cost, C, n_in, batch, batch_size_cfg, n_examples, 1,
                                        n_burn_in, adapt_forever, 1, epsilon, mdecay, scale_grad, seed, step,
                                        chain_offset, stream))
    return rc;
  const int sb = (int)(p->n_samples & 1);
  float* stage = p->d_stage + (int64_t)sb * C * p->D;
  if (host_sample != nullptr) {
    // the snapshot before last used this stage buffer: its D2H must have left the device
    if (p->n_samples >= 2) SG_CUDA(cudaStreamWaitEvent(main, p->sample_done[sb], 0));
    SG_CUDA(cudaMemcpyAsync(stage, theta, sizeof(float) * C * p->D, cudaMemcpyDeviceToDevice, main));
  }
  SG_CUDA(cudaEventRecord(p->step_done[b], main));
  SG_CUDA(cudaStreamWaitEvent(p->s_out, p->step_done[b], 0));
  SG_CUDA(cudaMemcpyAsync(host_cost, d_cost, sizeof(float) * C, cudaMemcpyDeviceToHost, p->s_out));
  if (host_sample != nullptr) {
    SG_CUDA(cudaMemcpyAsync(host_sample, stage, sizeof(float) * C * p->D, cudaMemcpyDeviceToHost, p->s_out));
    SG_CUDA(cudaEventRecord(p->sample_done[sb], p->s_out));
    ++p->n_samples;
  }
  SG_CUDA(cudaEventRecord(p->out_done[b], p->s_out));
  if (ticket != nullptr) *ticket = p->next_ticket;
  ++p->next_ticket;
  return SGMCMC_OK;
}

extern "C" int sgmcmc_bnn_host_pipeline_wait(sgmcmc_bnn_host_pipeline* p, int64_t ticket) {
  SG_REQUIRE(p != nullptr, SGMCMC_E_INVALID, "host_pipeline_wait: NULL handle");
  SG_REQUIRE(ticket >= 0 && ticket < p->next_ticket && ticket >= p->next_ticket - p->depth, SGMCMC_E_INVALID,
             "host_pipeline_wait: ticket %lld is not one of the last %d steps", (long long)ticket, p->depth);
  SG_CUDA(cudaEventSynchronize(p->out_done[ticket % p->depth]));
  return SGMCMC_OK;
}
