// Library-level plumbing of the C ABI: version, thread-local error string, launch
// counter and the launch-tuning knobs of the streaming kernels.
#include <atomic>
#include <stdarg.h>

#include "bnn_common.cuh"

namespace sgmcmc {

static thread_local char g_error[512] = "";
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_threads{256};
static std::atomic<int> g_unroll{1};
static std::atomic<int> g_update_max_ctas{0};
static std::atomic<int> g_update_carveout{-1};
static std::atomic<int> g_update_reverse{1};   // measured: -2 % (burn-in) / -3 % (sampling) per BNN step

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int check_launch(const char* what) {
  count_launch();
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(SGMCMC_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return SGMCMC_OK;
}

int tuning_threads() { return g_threads.load(std::memory_order_relaxed); }
int tuning_unroll() { return g_unroll.load(std::memory_order_relaxed); }
int tuning_update_max_ctas() { return g_update_max_ctas.load(std::memory_order_relaxed); }
int tuning_update_carveout() { return g_update_carveout.load(std::memory_order_relaxed); }
int tuning_update_reverse() { return g_update_reverse.load(std::memory_order_relaxed); }

}  // namespace sgmcmc

extern "C" {

int sgmcmc_version(void) { return 100; /* 0.1.0 */ }

const char* sgmcmc_last_error(void) { return sgmcmc::g_error; }

int64_t sgmcmc_launch_count(void) { return (int64_t)sgmcmc::g_launches.load(); }

int sgmcmc_set_update_tuning(int threads, int unroll) {
  if (threads != 0) {
    if (threads != 128 && threads != 256 && threads != 512)
      return sgmcmc::set_error(SGMCMC_E_INVALID, "threads must be 128, 256 or 512 (got %d)", threads);
    sgmcmc::g_threads.store(threads);
  }
  if (unroll != 0) {
    if (unroll != 1 && unroll != 2)
      return sgmcmc::set_error(SGMCMC_E_INVALID, "unroll must be 1 or 2 (got %d)", unroll);
    sgmcmc::g_unroll.store(unroll);
  }
  return SGMCMC_OK;
}

int sgmcmc_set_bnn_pipeline(int64_t chunk_chains, int ring) {
  if (chunk_chains < 0 || ring < 0) return sgmcmc::set_error(SGMCMC_E_INVALID, "pipeline: chunk and ring must be >= 0");
  sgmcmc::set_bnn_pipeline(chunk_chains, ring == 0 ? 2 : ring);
  // K1 must ask for the same shared-memory carveout as K4, or the two never share an SM
  sgmcmc::g_update_carveout.store(chunk_chains > 0 ? 100 : -1);
  return SGMCMC_OK;
}

int sgmcmc_set_update_reverse(int on) {
  sgmcmc::g_update_reverse.store(on & 3);   // bit 0: reverse walk, bit 1: theta stored with normal L2 priority
  return SGMCMC_OK;
}

int sgmcmc_set_persistent_grids(int update_max_ctas, int bnn_max_ctas) {
  if (update_max_ctas < 0 || bnn_max_ctas < 0)
    return sgmcmc::set_error(SGMCMC_E_INVALID, "grid caps must be >= 0");
  sgmcmc::g_update_max_ctas.store(update_max_ctas);
  sgmcmc::set_bnn_max_ctas(bnn_max_ctas);
  return SGMCMC_OK;
}

int sgmcmc_set_bnn_chunk(int64_t chains) {
  if (chains < 0) return sgmcmc::set_error(SGMCMC_E_INVALID, "chunk must be >= 0");
  sgmcmc::set_bnn_chunk(chains);
  return SGMCMC_OK;
}

int sgmcmc_set_bnn_fused(int on, int max_ctas) {
  if (max_ctas < 0) return sgmcmc::set_error(SGMCMC_E_INVALID, "max_ctas must be >= 0");
  sgmcmc::set_bnn_fused((on & 4) ? 2 : (on != 0 ? 1 : 0));
  sgmcmc::set_bnn_fused_prefetch((on & 2) == 0);
  sgmcmc::set_bnn_fused_max_ctas(max_ctas);
  return SGMCMC_OK;
}

int sgmcmc_set_bnn_tuning(int variant) {
  if (variant < 0 || variant >= sgmcmc::bnn_variant_count())
    return sgmcmc::set_error(SGMCMC_E_INVALID, "bnn variant must be in [0, %d) (got %d)",
                             sgmcmc::bnn_variant_count(), variant);
  sgmcmc::set_bnn_variant(variant);
  return SGMCMC_OK;
}

}  // extern "C"
