// Micro-benchmark: throughput of the legacy warp-level tensor path on sm_100a,
// mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32, with CH independent accumulator tiles
// per warp.  Reports cycles per mma per SM sub-partition and dense TFLOP/s.
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void k(float* out, int iters) {
  unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f800000u, 0x3f000000u, 0x3f800000u};
  unsigned b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
  float c[CH][4];
#pragma unroll
  for (int i = 0; i < CH; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
void run(int threads, int bps) {
  const int blocks = 148 * bps, iters = 1 << 13;
  float* out;
  cudaMalloc(&out, sizeof(float) * blocks * threads);
  k<CH><<<blocks, threads>>>(out, iters);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<CH><<<blocks, threads>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warps = (double)blocks * threads / 32;
  const double mmas = warps * iters * CH;
  const double cyc = ms * 1e-3 * 1.965e9;
  printf("CH=%d warps/SM=%3.0f: %.3f ms, %.2f cycles per mma per SMSP, %.1f dense TFLOP/s (tf32)\n", CH,
         warps / 148, ms, cyc / (mmas / (148.0 * 4)), 2.0 * 16 * 8 * 8 * mmas / ms / 1e9);
  cudaFree(out);
}

int main() {
  run<1>(128, 1); run<4>(128, 1); run<8>(128, 1);
  run<4>(256, 2); run<8>(256, 2); run<8>(512, 2);
  printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
