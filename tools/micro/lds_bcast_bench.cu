// Micro-benchmark: cost of warp-broadcast LDS.128 operand loads against FFMA work.
// Each thread does, per step: one 128-bit shared load from a warp-uniform address and
// FPL independent FFMA (FFMA-per-loaded-word = FPL / 4).  Reports cycles per step per SM
// partition so the shared-memory return bandwidth and the FFMA rate can be read off.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_bcast_bench lds_bcast_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int FPL, bool UNIFORM>
__global__ void k(float* out, int steps, int warps_stride) {
  __shared__ __align__(16) float buf[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) buf[i] = 1.0f + i * 1e-6f;
  __syncthreads();
  float acc[FPL];
#pragma unroll
  for (int f = 0; f < FPL; ++f) acc[f] = threadIdx.x + f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // UNIFORM: all lanes of a warp read the same 16 bytes; otherwise each lane its own 16 bytes
  const int base = UNIFORM ? warp * warps_stride : (warp * 32 + lane) * 4;
  float w[4] = {1.0001f, 0.9999f, 1.0002f, 0.9998f};
  for (int s = 0; s < steps; ++s) {
    const float4 h = *reinterpret_cast<const float4*>(&buf[(base + 4 * (s & 15)) & 4092]);
#pragma unroll
    for (int f = 0; f < FPL; f += 4) {
      acc[f + 0] = fmaf(h.x, w[0], acc[f + 0]);
      acc[f + 1] = fmaf(h.y, w[1], acc[f + 1]);
      acc[f + 2] = fmaf(h.z, w[2], acc[f + 2]);
      acc[f + 3] = fmaf(h.w, w[3], acc[f + 3]);
    }
  }
  float t = 0;
#pragma unroll
  for (int f = 0; f < FPL; ++f) t += acc[f];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int FPL, bool UNIFORM>
void run(const char* name, int threads, int blocks_per_sm) {
  const int blocks = 148 * blocks_per_sm, steps = 1 << 14;
  float* out;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  k<FPL, UNIFORM><<<blocks, threads>>>(out, steps, 16);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<FPL, UNIFORM><<<blocks, threads>>>(out, steps, 16);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warps_per_sm = threads / 32.0 * blocks_per_sm;
  const double cyc = ms * 1e-3 * 1.965e9;                       // SM cycles (max clock)
  const double lds_per_sm = warps_per_sm * steps;
  const double tflops = 2.0 * FPL * (double)blocks * threads * steps / ms / 1e9;
  printf("%-10s FPL=%2d warps/SM=%4.0f  %.3f ms  cycles per LDS.128 per SM = %.2f   FFMA %.1f TFLOP/s\n",
         name, FPL, warps_per_sm, ms, cyc / lds_per_sm, tflops);
  cudaFree(out);
}

int main() {
  for (int bps : {1, 2}) {
    run<4, true>("broadcast", 512, bps);
    run<8, true>("broadcast", 512, bps);
    run<16, true>("broadcast", 512, bps);
    run<32, true>("broadcast", 512, bps);
    run<4, false>("per-lane", 512, bps);
    run<8, false>("per-lane", 512, bps);
    run<16, false>("per-lane", 512, bps);
  }
  run<8, true>("broadcast", 128, 3);
  run<16, true>("broadcast", 128, 3);
  printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
