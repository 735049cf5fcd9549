// K11-K14: Stein variational gradient descent (pysgmcmc/samplers/svgd.py:81-182 and the
// pdist / squareform / median helpers of pysgmcmc/tensor_utils.py:160-208,326-577;
// restated in oracle/svgd.py).
//
// The particles ARE the engine's chain layout: X[n particles, D] fp32 row-major.  Unlike
// the other samplers the update couples all particles, so one step is
//   K11  squared distances         P[i,j] = (||x_i - x_j||)^2           n^2 D / 2 MAC
//   K12  svgd_select_*             median of the n^2 entries of P       exact radix select, 4 passes (L2)
//   K13  svgd_kernel_matrix_kernel K = exp(-P / h^2 / 2), row sums      n^2 exp, in place
//   K14  Stein direction           [K G | K X] (one GEMM, K read once) + AdaGrad history + update in
//                                  the epilogue                          4 n^2 D flop
// all enqueued on the caller's stream with no host round trip (the bandwidth h stays on the
// device).  K11 and K14 exist twice:
//   * here, on the FP32 pipe (FFMA; K11 subtracts before squaring like pdist): small or unaligned
//     shapes, and the second implementation the tests compare against;
//   * in svgd_sqdist_umma.cu / svgd_umma.cu on the tcgen05 tensor cores as 3xTF32 with TMEM
//     accumulators (fp32 accuracy; plain TF32 would cost 1e-3 relative in the Stein direction,
//     beyond the 1e-5 trajectory tolerance): 2.5-4x faster from a few hundred particles on.
// This file also holds the C entry points and the choice between the two.
#include <atomic>

#include "common.cuh"

namespace sgmcmc {

// ------------------------------------------------------------------------------------
// K11: squared pairwise distances.  64 x 64 output tile per CTA, 4 x 4 per thread; only
// tiles on or above the diagonal are computed, the mirror image goes out through a
// transposed shared-memory tile so both writes are coalesced.  Differences are formed
// before squaring (no Gram-matrix cancellation), and the result is (sqrt(s))^2 like the
// reference's squareform(pdist(.)) ** 2.
// ------------------------------------------------------------------------------------
constexpr int SD_T = 64;       // tile edge
constexpr int SD_BK = 16;      // feature chunk
constexpr int SD_LD = SD_T + 4;

__global__ void __launch_bounds__(256)
svgd_sqdist_kernel(const float* __restrict__ X, float* __restrict__ P, int n, int D) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  __shared__ __align__(16) float smem[SD_T * (SD_T + 1)];     // >= 2 * SD_BK * SD_LD
  float (*As)[SD_LD] = reinterpret_cast<float (*)[SD_LD]>(smem);
  float (*Bs)[SD_LD] = reinterpret_cast<float (*)[SD_LD]>(smem + SD_BK * SD_LD);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = bi * SD_T, j0 = bj * SD_T;
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;

  for (int k0 = 0; k0 < D; k0 += SD_BK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256, r = idx >> 4, k = idx & 15;
      const bool kin = (k0 + k) < D;
      As[k][r] = (kin && (i0 + r) < n) ? X[(int64_t)(i0 + r) * D + k0 + k] : 0.0f;
      Bs[k][r] = (kin && (j0 + r) < n) ? X[(int64_t)(j0 + r) * D + k0 + k] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SD_BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float d = av[r] - bv[c];
          acc[r][c] = fmaf(d, d, acc[r][c]);
        }
    }
    __syncthreads();
  }

#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float nrm = __fsqrt_rn(acc[r][c]);      // pdist: tf.norm(x_i - x_j)
      acc[r][c] = __fmul_rn(nrm, nrm);              // svgd.py:152: squareform(.) ** 2
    }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r;
    if (i < n) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = j0 + tx * 4 + c;
        if (j < n) P[(int64_t)i * n + j] = acc[r][c];
      }
    }
  }
  if (bi != bj) {
    // mirror image: transpose through shared memory (the k-loop ended with a barrier)
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) smem[(tx * 4 + c) * (SD_T + 1) + ty * 4 + r] = acc[r][c];
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int idx = tid + e * 256, jl = idx >> 6, il = idx & 63;
      if ((j0 + jl) < n && (i0 + il) < n) P[(int64_t)(j0 + jl) * n + i0 + il] = smem[jl * (SD_T + 1) + il];
    }
  }
}

// ------------------------------------------------------------------------------------
// K12: exact median by radix select on order-preserving 32-bit keys (4 passes of 8 bits).
// Two ranks are tracked at once (the two middle values of an even count).
// ------------------------------------------------------------------------------------
struct SelectState {
  unsigned long long rank[2];
  uint32_t prefix[2];
  uint32_t hist[2][256];
};
static_assert(sizeof(SelectState) <= 4096, "select scratch is documented as 4096 bytes");

__device__ __forceinline__ uint32_t float_key(uint32_t b) { return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u); }
__device__ __forceinline__ uint32_t key_float(uint32_t k) { return k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu); }

__global__ void svgd_select_init_kernel(SelectState* st, unsigned long long n_values) {
  const int t = threadIdx.x;
  st->hist[0][t] = 0;
  st->hist[1][t] = 0;
  if (t == 0) {
    // ascending 0-based ranks of the middle value(s) (tensor_utils.py:202-208)
    st->rank[1] = n_values / 2;
    st->rank[0] = (n_values % 2 == 1) ? n_values / 2 : n_values / 2 - 1;
    st->prefix[0] = 0;
    st->prefix[1] = 0;
  }
}

constexpr int SELECT_UNROLL = 8;

// Histogram increment for one element per lane.  The leading digit of distances clusters in two or three bins
// per warp, so pass 0 peels the distinct bins off with shuffle + ballot (one shared atomic per distinct bin);
// the later digits are spread out and go straight to shared-memory atomics.
__device__ __forceinline__ void warp_aggregated_inc(uint32_t* hist, uint32_t bin, bool active, bool clustered) {
  if (!clustered) {
    if (active) atomicAdd(&hist[bin], 1u);
    return;
  }
  unsigned remaining = __ballot_sync(0xFFFFFFFFu, active);
  const int lane = (int)(threadIdx.x & 31);
  while (remaining) {                                   // warp-uniform loop
    const int leader = __ffs(remaining) - 1;
    const uint32_t b = __shfl_sync(0xFFFFFFFFu, bin, leader);
    const unsigned same_bin = __ballot_sync(0xFFFFFFFFu, active && bin == b);
    if (lane == leader) atomicAdd(&hist[b], (uint32_t)__popc(same_bin));
    remaining &= ~same_bin;
  }
}

__global__ void __launch_bounds__(256)
svgd_select_hist_kernel(const uint32_t* __restrict__ values, int64_t n_values, int shift, SelectState* st) {
  __shared__ uint32_t h[2][256];
  h[0][threadIdx.x] = 0;
  h[1][threadIdx.x] = 0;
  __syncthreads();
  const uint32_t p0 = st->prefix[0], p1 = st->prefix[1];
  const uint32_t mask = (shift == 24) ? 0u : (0xFFFFFFFFu << (shift + 8));
  const bool same = (p0 == p1);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // all lanes of a warp run the same number of iterations (warp-synchronous helpers below)
  const int64_t n_round = (n_values + 31) / 32 * 32;
  // SELECT_UNROLL independent loads in flight per thread: the passes are bound by load latency otherwise
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_round; i0 += stride * SELECT_UNROLL) {
    uint32_t key[SELECT_UNROLL];
    bool in[SELECT_UNROLL];
#pragma unroll
    for (int u = 0; u < SELECT_UNROLL; ++u) {
      const int64_t i = i0 + u * stride;
      in[u] = i < n_values;
      key[u] = in[u] ? float_key(values[i]) : 0u;
    }
#pragma unroll
    for (int u = 0; u < SELECT_UNROLL; ++u) {
      const uint32_t bin = (key[u] >> shift) & 255u;
      warp_aggregated_inc(h[0], bin, in[u] && (key[u] & mask) == p0, shift == 24);
      if (!same) warp_aggregated_inc(h[1], bin, in[u] && (key[u] & mask) == p1, shift == 24);
    }
  }
  __syncthreads();
  if (h[0][threadIdx.x]) atomicAdd(&st->hist[0][threadIdx.x], h[0][threadIdx.x]);
  if (h[1][threadIdx.x]) atomicAdd(&st->hist[1][threadIdx.x], h[1][threadIdx.x]);
}

// The same histogram pass for a SYMMETRIC n x n matrix with a zero diagonal (the squared distances):
// only the upper triangle is read, every element counts twice and the n zeros of the diagonal are added
// by one thread -- half the traffic and half the atomics of the generic pass, same counts.
// Work items are (row, chunk of 256 * SELECT_UNROLL columns), dealt round-robin to a persistent grid.
__global__ void __launch_bounds__(256)
svgd_select_hist_sym_kernel(const uint32_t* __restrict__ values, int n, int shift, SelectState* st) {
  __shared__ uint32_t h[2][256];
  h[0][threadIdx.x] = 0;
  h[1][threadIdx.x] = 0;
  __syncthreads();
  const uint32_t p0 = st->prefix[0], p1 = st->prefix[1];
  const uint32_t mask = (shift == 24) ? 0u : (0xFFFFFFFFu << (shift + 8));
  const bool same = (p0 == p1);
  constexpr int CHUNK = 256 * SELECT_UNROLL;
  const int chunks = (n + CHUNK - 1) / CHUNK;
  const int64_t items = (int64_t)n * chunks;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int i = (int)(item / chunks), c0 = (int)(item - (int64_t)i * chunks) * CHUNK;
    if (c0 + CHUNK - 1 <= i) continue;                  // chunk entirely on or below the diagonal (block-uniform)
    const uint32_t* row = values + (int64_t)i * n;
    uint32_t key[SELECT_UNROLL];
    bool in[SELECT_UNROLL];
#pragma unroll
    for (int u = 0; u < SELECT_UNROLL; ++u) {
      const int j = c0 + u * 256 + (int)threadIdx.x;
      in[u] = j > i && j < n;
      key[u] = in[u] ? float_key(row[j]) : 0u;
    }
#pragma unroll
    for (int u = 0; u < SELECT_UNROLL; ++u) {
      const uint32_t bin = (key[u] >> shift) & 255u;
      warp_aggregated_inc(h[0], bin, in[u] && (key[u] & mask) == p0, shift == 24);
      if (!same) warp_aggregated_inc(h[1], bin, in[u] && (key[u] & mask) == p1, shift == 24);
    }
  }
  __syncthreads();
  // every off-diagonal value appears twice in the full matrix
  if (h[0][threadIdx.x]) atomicAdd(&st->hist[0][threadIdx.x], 2u * h[0][threadIdx.x]);
  if (h[1][threadIdx.x]) atomicAdd(&st->hist[1][threadIdx.x], 2u * h[1][threadIdx.x]);
  if (blockIdx.x == 0 && threadIdx.x == 0) {            // the n zeros of the diagonal
    const uint32_t key0 = float_key(0u), bin0 = (key0 >> shift) & 255u;
    if ((key0 & mask) == p0) atomicAdd(&st->hist[0][bin0], (uint32_t)n);
    if (!same && (key0 & mask) == p1) atomicAdd(&st->hist[1][bin0], (uint32_t)n);
  }
}

// Narrows both prefixes by 8 bits.  On the last pass writes out[0] = median and, when
// n_particles > 0, the RBF bandwidth of svgd.py:155-157: out[1] = h, out[2] = h^2.
__global__ void svgd_select_pick_kernel(SelectState* st, int shift, int last, float* out, float n_particles) {
  __shared__ uint32_t new_prefix[2];
  if (threadIdx.x < 64) {                               // warp 0 narrows rank 0, warp 1 rank 1 (counts < 2^31)
    const int r = (int)(threadIdx.x >> 5);
    const bool same = (st->prefix[0] == st->prefix[1]);
    const unsigned long long rank = st->rank[r];
    uint32_t bin, before;
    warp_pick_bin(same ? st->hist[0] : st->hist[r], (uint32_t)rank, bin, before);
    if ((threadIdx.x & 31) == 0) {
      new_prefix[r] = st->prefix[r] | (bin << shift);
      st->rank[r] = rank - before;
    }
  }
  __syncthreads();
  st->hist[0][threadIdx.x] = 0;
  st->hist[1][threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    st->prefix[0] = new_prefix[0];
    st->prefix[1] = new_prefix[1];
    if (last) {
      const float lo = __uint_as_float(key_float(new_prefix[0]));
      const float hi = __uint_as_float(key_float(new_prefix[1]));
      const float med = (new_prefix[0] == new_prefix[1]) ? lo : __fdiv_rn(__fadd_rn(hi, lo), 2.0f);
      out[0] = med;
      if (n_particles > 0.0f) {
        // h = sqrt(0.5 * median / log(n + 1))   (svgd.py:155-157)
        const float h = __fsqrt_rn(__fdiv_rn(__fmul_rn(0.5f, med), logf(__fadd_rn(n_particles, 1.0f))));
        out[1] = h;
        out[2] = __fmul_rn(h, h);
        out[3] = 0.0f;
      }
    }
  }
}

// Small inputs (up to SELECT_SMALL_MAX values, i.e. 362 particles): all four passes in ONE CTA with the
// histograms in shared memory -- one launch instead of nine; the result is the same exact selection.
constexpr int64_t SELECT_SMALL_MAX = 131072;

__global__ void __launch_bounds__(1024)
svgd_select_small_kernel(const uint32_t* __restrict__ values, int n_values, float* out, float n_particles) {
  __shared__ uint32_t h[2][256];
  __shared__ uint32_t prefix[2];
  __shared__ uint32_t rank[2];
  const int tid = threadIdx.x;
  if (tid == 0) {
    rank[1] = (uint32_t)(n_values / 2);
    rank[0] = (n_values % 2 == 1) ? (uint32_t)(n_values / 2) : (uint32_t)(n_values / 2 - 1);
    prefix[0] = prefix[1] = 0;
  }
  const int n_round = (n_values + 31) / 32 * 32;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    if (tid < 256) h[0][tid] = h[1][tid] = 0;
    __syncthreads();
    const uint32_t p0 = prefix[0], p1 = prefix[1];
    const uint32_t mask = (shift == 24) ? 0u : (0xFFFFFFFFu << (shift + 8));
    const bool same = (p0 == p1);
    for (int i0 = tid; i0 < n_round; i0 += 1024 * SELECT_UNROLL) {
      uint32_t key[SELECT_UNROLL];
      bool in[SELECT_UNROLL];
#pragma unroll
      for (int u = 0; u < SELECT_UNROLL; ++u) {
        const int i = i0 + u * 1024;
        in[u] = i < n_values;
        key[u] = in[u] ? float_key(values[i]) : 0u;
      }
#pragma unroll
      for (int u = 0; u < SELECT_UNROLL; ++u) {
        const uint32_t bin = (key[u] >> shift) & 255u;
        warp_aggregated_inc(h[0], bin, in[u] && (key[u] & mask) == p0, shift == 24);
        if (!same) warp_aggregated_inc(h[1], bin, in[u] && (key[u] & mask) == p1, shift == 24);
      }
    }
    __syncthreads();
    if (tid < 64) {                                     // warp 0 narrows rank 0, warp 1 rank 1
      const int w = tid >> 5;
      uint32_t bin, before;
      warp_pick_bin(same ? h[0] : h[w], rank[w], bin, before);
      if ((tid & 31) == 0) {
        prefix[w] |= bin << shift;
        rank[w] -= before;
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    const float lo = __uint_as_float(key_float(prefix[0]));
    const float hi = __uint_as_float(key_float(prefix[1]));
    const float med = (prefix[0] == prefix[1]) ? lo : __fdiv_rn(__fadd_rn(hi, lo), 2.0f);
    out[0] = med;
    if (n_particles > 0.0f) {
      const float hb = __fsqrt_rn(__fdiv_rn(__fmul_rn(0.5f, med), logf(__fadd_rn(n_particles, 1.0f))));
      out[1] = hb;
      out[2] = __fmul_rn(hb, hb);
      out[3] = 0.0f;
    }
  }
}

// symmetric_n > 0: `values` is a symmetric symmetric_n x symmetric_n matrix with a zero diagonal
static int launch_select(const float* values, int64_t n_values, float* out, void* scratch, float n_particles,
                         cudaStream_t stream, int symmetric_n = 0) {
  if (n_values <= SELECT_SMALL_MAX) {
    svgd_select_small_kernel<<<1, 1024, 0, stream>>>(reinterpret_cast<const uint32_t*>(values), (int)n_values, out,
                                                     n_particles);
    return check_launch("svgd_select_small_kernel");
  }
  SelectState* st = reinterpret_cast<SelectState*>(scratch);
  svgd_select_init_kernel<<<1, 256, 0, stream>>>(st, (unsigned long long)n_values);
  if (int rc = check_launch("svgd_select_init_kernel")) return rc;
  const int64_t want = (n_values + 256 * SELECT_UNROLL - 1) / (256 * SELECT_UNROLL);
  // 8 CTAs per SM: measured faster than 2 per SM (0.19 vs 0.25 ms at 16.7 M values) although every CTA ends
  // with up to 512 global atomics onto the same counters -- the passes want the parallelism
  const int grid = (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    if (symmetric_n > 0)
      svgd_select_hist_sym_kernel<<<148 * 8, 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(values), symmetric_n,
                                                               shift, st);
    else
      svgd_select_hist_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint32_t*>(values), n_values, shift, st);
    if (int rc = check_launch("svgd_select_hist_kernel")) return rc;
    svgd_select_pick_kernel<<<1, 256, 0, stream>>>(st, shift, pass == 3, out, n_particles);
    if (int rc = check_launch("svgd_select_pick_kernel")) return rc;
  }
  return SGMCMC_OK;
}

// ------------------------------------------------------------------------------------
// K13: K = exp(-P / h^2 / 2) in place, kernel_sum[i] = sum_j K[i,j]   (svgd.py:159-160)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
svgd_kernel_matrix_kernel(float* __restrict__ PK, float* __restrict__ ksum, const float* __restrict__ bw, int n) {
  __shared__ float red[8];
  const int row = blockIdx.x;
  const float h2 = bw[2];
  float s = 0.0f;
  float* p = PK + (int64_t)row * n;
  for (int j = threadIdx.x; j < n; j += 256) {
    const float k = expf(__fdiv_rn(__fdiv_rn(-p[j], h2), 2.0f));
    p[j] = k;
    s += k;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    ksum[row] = t;
  }
}

// ------------------------------------------------------------------------------------
// K14: Stein direction + AdaGrad + update.  C tile 128 (particles i) x 64 (dims d),
// contraction over the particles j in chunks of 16; two accumulators per output share
// the K operand: KG = K @ grad and KX = K @ X.  K is symmetric (bit-wise: K11 mirrors its
// tiles), so the A tile is read as rows of K -- contiguous along i, no transpose.
// 256 threads, 8 x 4 outputs per thread: per k-step 4 LDS.128 feed 64 FFMA.
// Epilogue (svgd.py:130-148), op by op in the reference's order:
//   kgrad = (-KX + x * ksum_i) / h^2 ;  phi = (KG + kgrad) / n
//   hist  = alpha * hist + (1 - alpha) * phi^2 ;  x_new = x - eps * phi / (fudge + sqrt(hist))
// x_new goes to a second buffer: every CTA reads all of X.
// ------------------------------------------------------------------------------------
constexpr int SU_BM = 128, SU_BN = 64, SU_BK = 16;

template <bool VEC>
struct SvgdTileLoader {
  float4 a[2], g, x;     // VEC: two float4 of K, one of grad, one of X per thread
  float as[8], gs[4], xs[4];

  __device__ __forceinline__ void load(const float* __restrict__ K, const float* __restrict__ G,
                                       const float* __restrict__ X, int n, int D, int i0, int d0, int k0, int tid) {
    if (VEC) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int idx = tid + e * 256, k = idx >> 5, i4 = (idx & 31) * 4;
        a[e] = ((k0 + k) < n && (i0 + i4) < n)
                   ? *reinterpret_cast<const float4*>(K + (int64_t)(k0 + k) * n + i0 + i4)
                   : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const int k = tid >> 4, d4 = (tid & 15) * 4;
      const bool in = (k0 + k) < n && (d0 + d4) < D;
      const int64_t off = (int64_t)(k0 + k) * D + d0 + d4;
      g = in ? *reinterpret_cast<const float4*>(G + off) : make_float4(0.f, 0.f, 0.f, 0.f);
      x = in ? *reinterpret_cast<const float4*>(X + off) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int idx = tid + e * 256, k = idx >> 7, i = idx & 127;
        as[e] = ((k0 + k) < n && (i0 + i) < n) ? K[(int64_t)(k0 + k) * n + i0 + i] : 0.0f;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = tid + e * 256, k = idx >> 6, d = idx & 63;
        const bool in = (k0 + k) < n && (d0 + d) < D;
        const int64_t off = (int64_t)(k0 + k) * D + d0 + d;
        gs[e] = in ? G[off] : 0.0f;
        xs[e] = in ? X[off] : 0.0f;
      }
    }
  }

  __device__ __forceinline__ void store(float (*As)[SU_BM], float (*Bg)[SU_BN], float (*Bx)[SU_BN], int tid) const {
    if (VEC) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int idx = tid + e * 256, k = idx >> 5, i4 = (idx & 31) * 4;
        *reinterpret_cast<float4*>(&As[k][i4]) = a[e];
      }
      const int k = tid >> 4, d4 = (tid & 15) * 4;
      *reinterpret_cast<float4*>(&Bg[k][d4]) = g;
      *reinterpret_cast<float4*>(&Bx[k][d4]) = x;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int idx = tid + e * 256;
        As[idx >> 7][idx & 127] = as[e];
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = tid + e * 256;
        Bg[idx >> 6][idx & 63] = gs[e];
        Bx[idx >> 6][idx & 63] = xs[e];
      }
    }
  }
};

template <bool VEC>
__global__ void __launch_bounds__(256, 2)
svgd_update_kernel(const float* __restrict__ K, const float* __restrict__ X, const float* __restrict__ G,
                   const float* __restrict__ ksum, const float* __restrict__ bw, float* __restrict__ hist,
                   float* __restrict__ Xout, int n, int D, float eps, float alpha, float one_minus_alpha,
                   float fudge) {
  __shared__ __align__(16) float As[2][SU_BK][SU_BM];
  __shared__ __align__(16) float Bg[2][SU_BK][SU_BN];
  __shared__ __align__(16) float Bx[2][SU_BK][SU_BN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int d0 = blockIdx.x * SU_BN, i0 = blockIdx.y * SU_BM;

  float accg[8][4], accx[8][4];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) accg[r][c] = accx[r][c] = 0.0f;

  SvgdTileLoader<VEC> ld;
  const int n_tiles = (n + SU_BK - 1) / SU_BK;
  ld.load(K, G, X, n, D, i0, d0, 0, tid);
  ld.store(As[0], Bg[0], Bx[0], tid);
  __syncthreads();
  for (int t = 0; t < n_tiles; ++t) {
    const int cur = t & 1;
    if (t + 1 < n_tiles) ld.load(K, G, X, n, D, i0, d0, (t + 1) * SU_BK, tid);
#pragma unroll
    for (int k = 0; k < SU_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      const float4 bg = *reinterpret_cast<const float4*>(&Bg[cur][k][tx * 4]);
      const float4 bx = *reinterpret_cast<const float4*>(&Bx[cur][k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float gv[4] = {bg.x, bg.y, bg.z, bg.w}, xv[4] = {bx.x, bx.y, bx.z, bx.w};
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          accg[r][c] = fmaf(av[r], gv[c], accg[r][c]);
          accx[r][c] = fmaf(av[r], xv[c], accx[r][c]);
        }
    }
    if (t + 1 < n_tiles) ld.store(As[cur ^ 1], Bg[cur ^ 1], Bx[cur ^ 1], tid);
    __syncthreads();
  }

  const float h2 = bw[2];
  const float nf = (float)n;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int i = i0 + (r < 4 ? ty * 4 + r : 64 + ty * 4 + (r - 4));
    if (i >= n) continue;
    const float ks = ksum[i];
    const int d = d0 + tx * 4;
    const int64_t off = (int64_t)i * D + d;
    float xv[4], hv[4];
    if (VEC) {
      if (d >= D) continue;
      const float4 x4 = *reinterpret_cast<const float4*>(X + off);
      const float4 h4 = *reinterpret_cast<const float4*>(hist + off);
      xv[0] = x4.x; xv[1] = x4.y; xv[2] = x4.z; xv[3] = x4.w;
      hv[0] = h4.x; hv[1] = h4.y; hv[2] = h4.z; hv[3] = h4.w;
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        xv[c] = (d + c) < D ? X[off + c] : 0.0f;
        hv[c] = (d + c) < D ? hist[off + c] : 0.0f;
      }
    }
    float xo[4], ho[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float kgrad = __fdiv_rn(__fadd_rn(-accx[r][c], __fmul_rn(xv[c], ks)), h2);
      const float phi = __fdiv_rn(__fadd_rn(accg[r][c], kgrad), nf);
      ho[c] = __fadd_rn(__fmul_rn(alpha, hv[c]), __fmul_rn(one_minus_alpha, __fmul_rn(phi, phi)));
      const float adj = __fdiv_rn(phi, __fadd_rn(fudge, __fsqrt_rn(ho[c])));
      xo[c] = __fsub_rn(xv[c], __fmul_rn(eps, adj));
    }
    if (VEC) {
      *reinterpret_cast<float4*>(hist + off) = make_float4(ho[0], ho[1], ho[2], ho[3]);
      *reinterpret_cast<float4*>(Xout + off) = make_float4(xo[0], xo[1], xo[2], xo[3]);
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if ((d + c) < D) {
          hist[off + c] = ho[c];
          Xout[off + c] = xo[c];
        }
    }
  }
}

// csrc/svgd_umma.cu: the same update on the tcgen05 tensor cores (3xTF32, TMEM accumulators)
int launch_svgd_update_umma(const float* K, const float* X, const float* G, const float* ksum, const float* bw,
                            float* hist, float* Xout, int n, int D, float eps, float alpha, float one_minus_alpha,
                            float fudge, int prefetch, cudaStream_t stream);

// 0 = auto, 1 = FFMA kernel, 2 = tcgen05 kernel wherever it is eligible; 22/23/24 = tcgen05 with 2/3/4
// producer register buffers (sweeps)
static std::atomic<int> g_svgd_impl{0};

// csrc/svgd_sqdist_umma.cu: K11 through the Gram matrix of the centred particles on the tensor cores
int launch_svgd_sqdist_umma(const float* X, float* P, float* work, int64_t work_floats, int n, int D,
                            cudaStream_t stream);
int64_t svgd_sqdist_work_floats(int n, int D, int n_slices);
int svgd_sqdist_best_slices(int n, int D);

static int check_svgd_sizes(int64_t n, int64_t D) {
  SG_REQUIRE(n >= 0 && D >= 0, SGMCMC_E_INVALID, "svgd: n_particles and n_dims must be >= 0");
  SG_REQUIRE(n <= 46340, SGMCMC_E_UNSUPPORTED, "svgd: at most 46340 particles (n^2 must fit 31 bits), got %lld", (long long)n);
  SG_REQUIRE(D <= (int64_t)1 << 30 && n * D < ((int64_t)1 << 40), SGMCMC_E_UNSUPPORTED, "svgd: n_dims too large");
  return SGMCMC_OK;
}

}  // namespace sgmcmc

using namespace sgmcmc;

extern "C" int sgmcmc_set_svgd_tuning(int impl) {
  SG_REQUIRE((impl >= 0 && impl <= 2) || (impl >= 22 && impl <= 24), SGMCMC_E_INVALID,  /* 22 == 2 */
             "svgd impl must be 0 (auto), 1 (FFMA), 2 (tcgen05) or 22-24 (tcgen05, 2-4 producer buffers), got %d", impl);
  g_svgd_impl.store(impl);
  return SGMCMC_OK;
}

extern "C" int sgmcmc_median_f32(const float* values, int64_t n_values, float* out, void* scratch, void* stream) {
  SG_REQUIRE(n_values >= 1, SGMCMC_E_INVALID, "median: needs at least one value");
  SG_REQUIRE(n_values < ((int64_t)1 << 31), SGMCMC_E_UNSUPPORTED, "median: at most 2^31 - 1 values");
  SG_REQUIRE(values && out && scratch, SGMCMC_E_INVALID, "median: NULL pointer");
  SG_REQUIRE(aligned_to(values, 4) && aligned_to(out, 4) && aligned_to(scratch, 8), SGMCMC_E_ALIGN, "median: misaligned pointer");
  return launch_select(values, n_values, out, scratch, 0.0f, (cudaStream_t)stream);
}

extern "C" int sgmcmc_median_symmetric_f32(const float* matrix, int64_t n, float* out, void* scratch, void* stream) {
  SG_REQUIRE(n >= 1 && n <= 46340, SGMCMC_E_INVALID, "median: n must be in [1, 46340]");
  SG_REQUIRE(matrix && out && scratch, SGMCMC_E_INVALID, "median: NULL pointer");
  SG_REQUIRE(aligned_to(matrix, 4) && aligned_to(out, 4) && aligned_to(scratch, 8), SGMCMC_E_ALIGN, "median: misaligned pointer");
  return launch_select(matrix, n * n, out, scratch, 0.0f, (cudaStream_t)stream, (int)n);
}

extern "C" int64_t sgmcmc_svgd_scratch_bytes(int64_t n_particles, int64_t n_dims) {
  if (n_particles < 0 || n_dims < 0 || n_particles > 46340 || n_dims > ((int64_t)1 << 30)) return -1;
  const int n = (int)n_particles, D = (int)n_dims;
  return 4096 + 4 * svgd_sqdist_work_floats(n, D, (n >= 1 && D >= 1) ? svgd_sqdist_best_slices(n, D) : 1);
}

extern "C" int sgmcmc_svgd_kernel_matrix_f32(const float* particles, float* kernel_matrix, float* kernel_sum,
                                             float* bandwidth, void* scratch, int64_t scratch_bytes,
                                             int64_t n_particles, int64_t n_dims, void* stream) {
  if (int rc = check_svgd_sizes(n_particles, n_dims)) return rc;
  SG_REQUIRE(n_particles >= 1 && n_dims >= 1, SGMCMC_E_INVALID, "svgd: needs at least one particle and one dimension");
  SG_REQUIRE(particles && kernel_matrix && kernel_sum && bandwidth && scratch, SGMCMC_E_INVALID, "svgd: NULL pointer");
  SG_REQUIRE(scratch_bytes >= 4096, SGMCMC_E_INVALID, "svgd: scratch must hold at least 4096 bytes");
  SG_REQUIRE(aligned_to(particles, 4) && aligned_to(kernel_matrix, 4) && aligned_to(kernel_sum, 4) &&
                 aligned_to(bandwidth, 4) && aligned_to(scratch, 8),
             SGMCMC_E_ALIGN, "svgd: misaligned pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const int n = (int)n_particles, D = (int)n_dims;
  const int impl = g_svgd_impl.load(std::memory_order_relaxed);
  const int64_t work_floats = (scratch_bytes - 4096) / 4;
  const bool umma_ok = (D % 4 == 0) && aligned_to(particles, 16) && aligned_to(scratch, 16) &&
                       work_floats >= svgd_sqdist_work_floats(n, D, 1);
  const bool use_umma = impl >= 2 ? umma_ok : (impl == 0 && umma_ok && n >= 256 && D >= 128);
  if (use_umma) {
    float* work = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(scratch) + 4096);
    if (int rc = launch_svgd_sqdist_umma(particles, kernel_matrix, work, work_floats, n, D, s)) return rc;
  } else {
    const unsigned nt = (unsigned)((n + SD_T - 1) / SD_T);
    svgd_sqdist_kernel<<<dim3(nt, nt), 256, 0, s>>>(particles, kernel_matrix, n, D);
    if (int rc = check_launch("svgd_sqdist_kernel")) return rc;
  }
  if (int rc = launch_select(kernel_matrix, n_particles * n_particles, bandwidth, scratch, (float)n, s, n)) return rc;
  svgd_kernel_matrix_kernel<<<n, 256, 0, s>>>(kernel_matrix, kernel_sum, bandwidth, n);
  return check_launch("svgd_kernel_matrix_kernel");
}

extern "C" int sgmcmc_svgd_update_f32(float* particles, const float* grad, float* historical_grad,
                                      const float* kernel_matrix, const float* kernel_sum, const float* bandwidth,
                                      float* particles_scratch, int64_t n_particles, int64_t n_dims, float epsilon,
                                      float alpha, float one_minus_alpha, float fudge_factor, void* stream) {
  if (int rc = check_svgd_sizes(n_particles, n_dims)) return rc;
  if (n_particles == 0 || n_dims == 0) return SGMCMC_OK;
  SG_REQUIRE(particles && grad && historical_grad && kernel_matrix && kernel_sum && bandwidth && particles_scratch,
             SGMCMC_E_INVALID, "svgd: NULL pointer");
  SG_REQUIRE(aligned_to(particles, 4) && aligned_to(grad, 4) && aligned_to(historical_grad, 4) &&
                 aligned_to(kernel_matrix, 4) && aligned_to(kernel_sum, 4) && aligned_to(particles_scratch, 4),
             SGMCMC_E_ALIGN, "svgd: misaligned pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const int n = (int)n_particles, D = (int)n_dims;
  const dim3 grid((unsigned)((D + SU_BN - 1) / SU_BN), (unsigned)((n + SU_BM - 1) / SU_BM));
  SG_REQUIRE(grid.y <= 65535, SGMCMC_E_UNSUPPORTED, "svgd: too many particles");
  const bool vec = (n % 4 == 0) && (D % 4 == 0) && aligned_to(particles, 16) && aligned_to(grad, 16) &&
                   aligned_to(historical_grad, 16) && aligned_to(kernel_matrix, 16) &&
                   aligned_to(particles_scratch, 16);
  const int impl = g_svgd_impl.load(std::memory_order_relaxed);
  const bool umma_ok = vec && aligned_to(kernel_sum, 4);
  // auto: the tensor-core kernel needs enough rows to fill its 128-row tile and enough work to amortise its
  // prologue; measured (profiles/r01_svgd_tcgen05.jsonl): 128 x 128 FFMA 0.029 vs 0.032 ms, 256 x 5252 0.070 vs
  // 0.044 ms, 4096 x 64 0.50 vs 0.23 ms
  const bool use_umma = impl >= 2 ? umma_ok
                                  : (impl == 0 && umma_ok && n >= 128 && D >= 64 && (int64_t)n * D >= 65536);
  if (use_umma) {
    if (int rc = launch_svgd_update_umma(kernel_matrix, particles, grad, kernel_sum, bandwidth, historical_grad,
                                         particles_scratch, n, D, epsilon, alpha, one_minus_alpha, fudge_factor,
                                         impl >= 20 ? impl - 20 : 0, s))
      return rc;
  } else if (vec)
    svgd_update_kernel<true><<<grid, 256, 0, s>>>(kernel_matrix, particles, grad, kernel_sum, bandwidth,
                                                  historical_grad, particles_scratch, n, D, epsilon, alpha,
                                                  one_minus_alpha, fudge_factor);
  else
    svgd_update_kernel<false><<<grid, 256, 0, s>>>(kernel_matrix, particles, grad, kernel_sum, bandwidth,
                                                   historical_grad, particles_scratch, n, D, epsilon, alpha,
                                                   one_minus_alpha, fudge_factor);
  if (!use_umma)
    if (int rc = check_launch("svgd_update_kernel")) return rc;
  const cudaError_t e = cudaMemcpyAsync(particles, particles_scratch, sizeof(float) * (size_t)n * (size_t)D,
                                        cudaMemcpyDeviceToDevice, s);
  if (e != cudaSuccess) return set_error(SGMCMC_E_CUDA, "svgd: cudaMemcpyAsync: %s", cudaGetErrorString(e));
  return SGMCMC_OK;
}
