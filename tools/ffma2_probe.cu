// FP32 FMA issue rates on this GPU: scalar FFMA vs packed FFMA2 (fma.rn.f32x2), multiplier pair from
// vector registers or straight from kernel parameters (uniform registers, as the convolution kernels do).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
  unsigned long long da = *reinterpret_cast<unsigned long long*>(&d);
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(da) : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  d = *reinterpret_cast<float2*>(&da);
}
struct K8 { float2 kk[8]; };
template <int MODE>
__global__ void k(float* out, int iters, float s, K8 kp) {
  float2 acc[8];
  float2 a = make_float2(threadIdx.x * 1e-3f, s), b = make_float2(s, s * 0.5f);
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(i, -i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) { acc[i].x = fmaf(a.x, b.x, acc[i].x); acc[i].y = fmaf(a.y, b.y, acc[i].y); }
        else if (MODE == 1) ffma2(acc[i], a, b);
        else if (MODE == 2) ffma2(acc[i], a, kp.kk[u]);                       // multiplier pair from parameters
        else { acc[i].x = fmaf(a.x, kp.kk[u].x, acc[i].x); acc[i].y = fmaf(a.y, kp.kk[u].y, acc[i].y); }
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
  K8 kp; for (int i = 0; i < 8; ++i) kp.kk[i] = make_float2(1e-6f * (i + 1), 1e-6f * (i + 1));
  const char* names[4] = {"FFMA  (registers)", "FFMA2 (registers)", "FFMA2 (uniform/param multiplier)", "FFMA  (uniform/param multiplier)"};
  for (int mode = 0; mode < 4; ++mode) for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    if (mode == 0) k<0><<<148 * 4, 256>>>(d, iters, 1e-6f, kp);
    else if (mode == 1) k<1><<<148 * 4, 256>>>(d, iters, 1e-6f, kp);
    else if (mode == 2) k<2><<<148 * 4, 256>>>(d, iters, 1e-6f, kp);
    else k<3><<<148 * 4, 256>>>(d, iters, 1e-6f, kp);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = 148.0 * 4 * 256 * iters * 64 * 2;
    if (rep == 2) printf("%-34s %.3f ms, %.2f T FMA/s, %.1f FMA/clk/SM at 1.92 GHz\n", names[mode], ms, fmas / ms / 1e9, fmas / (ms * 1e-3) / 148 / 1.92e9);
  }
  return 0;
}
