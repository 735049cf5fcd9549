#include "common.cuh"
using namespace sgmcmc;
extern "C" int sgmcmc_chain_moments_f32(const float*, double*, int64_t, int64_t, int64_t, void*) { return set_error(SGMCMC_E_UNSUPPORTED, "not built yet"); }
extern "C" int sgmcmc_variogram_f32(const float*, double*, int64_t, int64_t, int64_t, int64_t, int64_t, void*) { return set_error(SGMCMC_E_UNSUPPORTED, "not built yet"); }
