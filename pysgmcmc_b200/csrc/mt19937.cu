// K7: on-device minibatch start indices, bit-exact with the reference's
//     rng = numpy.random.RandomState(seed); start = rng.randint(0, N - B + 1)
// (pysgmcmc/data_batches.py:104-120).  numpy's legacy generator is MT19937 seeded with
// init_genrand(seed) and its bounded integer is "AND with the smallest all-ones mask
// >= max, reject while > max" on successive 32-bit outputs (oracle/mt19937.py, pinned
// against numpy itself).
//
// One thread owns one stream (= one chain).  The 624-word state lives in global memory
// stream-minor ([625, n_streams]: word i of stream j at state[i*n_streams + j], row 624 =
// position), so the per-thread sequential walks of the twist are coalesced across the
// warp.  MT19937 is inherently sequential per stream and the rejection loop is data
// dependent; the kernel is latency bound and tiny next to the BNN step (one twist per
// ~380 steps at N-B = 19980).
#include "common.cuh"

namespace sgmcmc {

constexpr int MT_N = 624, MT_M = 397;

__global__ void mt19937_seed_kernel(uint32_t* __restrict__ state, const uint32_t* __restrict__ seeds,
                                    int64_t n_streams) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_streams) return;
  uint32_t x = seeds[j];
  state[j] = x;
  for (int i = 1; i < MT_N; ++i) {
    x = 1812433253u * (x ^ (x >> 30)) + (uint32_t)i;
    state[(int64_t)i * n_streams + j] = x;
  }
  state[(int64_t)MT_N * n_streams + j] = MT_N;   // position: forces a twist on first use
}

__device__ __forceinline__ uint32_t mt_mix(uint32_t cur, uint32_t nxt, uint32_t far) {
  const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
  return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

// The twist of ONE stream (mt[i] = mt[(i + 397) % 624] ^ f(mt[i], mt[(i + 1) % 624]), in place, ascending i) done
// by a whole warp, 32 consecutive i per round: within a round every lane reads its three words before any lane
// writes (mt[i + 1] is the neighbour's old word), and a round only depends on words written >= 195 positions
// earlier (i - 227) or not yet written (i + 1, i + 397), so 20 rounds replace 624 dependent iterations.  Lanes
// of a warp consume their streams at different rates (rejection sampling), so they reach a twist at different
// steps: done one lane at a time, sequentially, a warp would walk the 624-iteration loop up to 32 times.
__device__ void mt_twist_warp(uint32_t* __restrict__ st, int64_t n_streams, int64_t j, int lane) {
  for (int i0 = 0; i0 < MT_N; i0 += 32) {
    const int i = i0 + lane;
    uint32_t nv = 0;
    if (i < MT_N) {
      const int i1 = i + 1 < MT_N ? i + 1 : 0;
      const int k = i + MT_M < MT_N ? i + MT_M : i + MT_M - MT_N;
      nv = mt_mix(st[(int64_t)i * n_streams + j], st[(int64_t)i1 * n_streams + j], st[(int64_t)k * n_streams + j]);
    }
    __syncwarp();
    if (i < MT_N) st[(int64_t)i * n_streams + j] = nv;
    __syncwarp();
  }
}

__global__ void mt19937_starts_kernel(uint32_t* __restrict__ state, int32_t* __restrict__ starts,
                                      int64_t n_streams, int64_t n_steps, uint32_t max_inclusive,
                                      uint32_t mask) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool valid = j < n_streams;                   // whole warps stay: the twist is a warp-wide operation
  uint32_t pos = valid ? state[(int64_t)MT_N * n_streams + j] : 0;
  for (int64_t s = 0; s < n_steps; ++s) {
    uint32_t v = 0;
    bool done = !valid || max_inclusive == 0;         // max == 0 consumes nothing (numpy legacy behaviour)
    while (!__all_sync(0xffffffffu, done)) {
      unsigned need = __ballot_sync(0xffffffffu, !done && pos >= MT_N);
      while (need != 0) {
        const int src = __ffs(need) - 1;
        need &= need - 1;
        mt_twist_warp(state, n_streams, __shfl_sync(0xffffffffu, j, src), lane);
        if (lane == src) pos = 0;
      }
      if (!done) {
        uint32_t y = state[(int64_t)pos * n_streams + j];
        ++pos;
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        v = y & mask;
        done = v <= max_inclusive;
      }
    }
    if (valid) starts[s * n_streams + j] = (int32_t)v;
  }
  if (valid) state[(int64_t)MT_N * n_streams + j] = pos;
}

}  // namespace sgmcmc

using namespace sgmcmc;

extern "C" int sgmcmc_mt19937_seed(uint32_t* state, const uint32_t* seeds, int64_t n_streams, void* stream) {
  SG_REQUIRE(n_streams >= 0, SGMCMC_E_INVALID, "n_streams must be >= 0");
  if (n_streams == 0) return SGMCMC_OK;
  SG_REQUIRE(state && seeds, SGMCMC_E_INVALID, "mt19937_seed: NULL pointer");
  const int threads = 128;
  mt19937_seed_kernel<<<(unsigned)((n_streams + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      state, seeds, n_streams);
  return check_launch("mt19937_seed_kernel");
}

extern "C" int sgmcmc_mt19937_starts(uint32_t* state, int32_t* starts, int64_t n_streams, int64_t n_steps,
                                     uint32_t max_inclusive, void* stream) {
  SG_REQUIRE(n_streams >= 0 && n_steps >= 0, SGMCMC_E_INVALID, "negative size");
  if (n_streams == 0 || n_steps == 0) return SGMCMC_OK;
  SG_REQUIRE(state && starts, SGMCMC_E_INVALID, "mt19937_starts: NULL pointer");
  SG_REQUIRE(max_inclusive <= 0x7fffffffu, SGMCMC_E_INVALID, "max_inclusive must fit int32");
  uint32_t mask = max_inclusive;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  const int threads = 128;
  mt19937_starts_kernel<<<(unsigned)((n_streams + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
      state, starts, n_streams, n_steps, max_inclusive, mask);
  return check_launch("mt19937_starts_kernel");
}
