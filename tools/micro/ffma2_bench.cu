// Micro-benchmark: FP32 FMA throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
  unsigned long long rc = *reinterpret_cast<unsigned long long*>(&c);
  unsigned long long rd;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

template <int CH>
__global__ void k_ffma(float* out, int iters, float a, float b) {
  float acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = threadIdx.x * 1e-3f + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = fmaf(acc[c], a, b);
  }
  float s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void k_ffma2(float* out, int iters, float a, float b) {
  float2 acc[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) acc[c] = make_float2(threadIdx.x * 1e-3f + c, c);
  const float2 a2 = make_float2(a, a * 0.5f), b2 = make_float2(b, b * 2.0f);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = ffma2(acc[c], a2, b2);
  }
  float s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += acc[c].x + acc[c].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  const int blocks = 148 * 8, threads = 256, iters = 1 << 14;
  float* out;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  constexpr int CH = 8;
  const double fma1 = (double)blocks * threads * iters * CH;
  float ms1 = time_ms([&] { k_ffma<CH><<<blocks, threads>>>(out, iters, 0.999f, 0.001f); });
  float ms2 = time_ms([&] { k_ffma2<CH><<<blocks, threads>>>(out, iters, 0.999f, 0.001f); });
  printf("FFMA : %.3f ms  %.1f TFLOP/s\n", ms1, 2 * fma1 / ms1 / 1e9);
  printf("FFMA2: %.3f ms  %.1f TFLOP/s (2 FMA per instruction)\n", ms2, 4 * fma1 / ms2 / 1e9);
  printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
