#include "common.cuh"
using namespace sgmcmc;
extern "C" int sgmcmc_bnn_nll_grad_f32(const float*, const float*, const float*, const int32_t*, float*, float*, float*, int64_t, int, int, float, int64_t, void*) { return set_error(SGMCMC_E_UNSUPPORTED, "not built yet"); }
extern "C" int sgmcmc_bnn_sghmc_run_f32(float*, float*, float*, float*, float*, float*, const float*, const float*, const int32_t*, const float*, float*, float*, int64_t, int, int, float, int64_t, int64_t, int64_t, int, int64_t, float, float, float, uint64_t, uint64_t, uint64_t, void*) { return set_error(SGMCMC_E_UNSUPPORTED, "not built yet"); }
extern "C" int sgmcmc_bnn_predict_f32(const float*, const float*, float*, int64_t, int, int64_t, void*) { return set_error(SGMCMC_E_UNSUPPORTED, "not built yet"); }
