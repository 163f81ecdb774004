#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
  unsigned long long da = *reinterpret_cast<unsigned long long*>(&d);
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(da) : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  d = *reinterpret_cast<float2*>(&da);
}
template <int MODE>
__global__ void k(float* out, int iters, float s) {
  float2 acc[8];
  float2 a = make_float2(threadIdx.x * 1e-3f, s), b = make_float2(s, s * 0.5f);
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(i, -i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) { acc[i].x = fmaf(a.x, b.x, acc[i].x); acc[i].y = fmaf(a.y, b.y, acc[i].y); }
        else ffma2(acc[i], a, b);
      }
    }
  }
  float r = 0; for (int i = 0; i < 8; ++i) r += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4096;
  for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    if (mode == 0) k<0><<<148 * 4, 256>>>(d, iters, 1e-6f); else k<1><<<148 * 4, 256>>>(d, iters, 1e-6f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = 148.0 * 4 * 256 * iters * 64 * 2;
    printf("mode %d: %.3f ms, %.2f TFMA/s (%.1f fma/clk/SM at 1.9GHz)\n", mode, ms, fmas / ms / 1e9, fmas / ms / 1e3 / 148 / 1.9e6);
  }
  return 0;
}
