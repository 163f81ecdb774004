// What does HBM give a kernel that mostly WRITES?  l0_fused_kernel reads 1 B and writes 12 B per
// pixel (108 MB per 4K frame, 92 % of it stores); the roofline denominator (MEASURED_PEAKS.json) is
// a copy, i.e. half reads, half writes.  This probe times, back to back between one pair of events
// and on buffers that together exceed the 126 MB L2:
//   fill      : 3 planes of 3840x2160 floats written with STG.128 (99.5 MB / launch), nothing read
//   fill+u8   : the same plus the 8.3 MB u8 frame read (the level-0 kernel's exact traffic)
//   copy      : float4 copy of one plane (33 MB -> 33 MB, half reads)
//   memset    : cudaMemsetAsync of the 3 planes
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/write_probe.cu -o tools/write_probe.bin
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void fill3(float4* a, float4* b, float4* c, size_t n4, float v) {
  const float4 w = make_float4(v, v + 1, v + 2, v + 3);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    a[i] = w; b[i] = w; c[i] = w;
  }
}
__global__ void fill3_u8(const uchar4* __restrict__ src, float4* a, float4* b, float4* c, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const uchar4 p = src[i];
    const float4 w = make_float4(p.x, p.y, p.z, p.w);
    a[i] = w; b[i] = make_float4(w.y, w.z, w.w, w.x); c[i] = make_float4(w.z, w.w, w.x, w.y);
  }
}
__global__ void copy4(const float4* __restrict__ s, float4* d, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
}

// the tile kernels' store pattern: persistent CTAs claim TW x TH tiles in raster order; a warp writes
// 32 / (TW / 4) row segments of TW * 4 bytes per instruction, plane after plane
template <int TW, int TH>
__global__ void fill3_tiled(float* a, float* b, float* c, int W, int H, unsigned* ctr, unsigned base, const unsigned char* u8) {
  __shared__ int s_tile;
  const int tiles_x = W / TW, ntiles = tiles_x * (H / TH);
  constexpr int CG = TW / 4, RPP = 256 / CG;            // column groups, rows per pass
  const int j = threadIdx.x % CG, r0 = threadIdx.x / CG;
  for (;;) {
    if (threadIdx.x == 0) s_tile = (int)(atomicAdd(ctr, 1u) - base);
    __syncthreads();
    const int t = s_tile;
    __syncthreads();
    if (t >= ntiles) break;
    const int x0 = (t % tiles_x) * TW, y0 = (t / tiles_x) * TH;
    float v = 1.0f;
    if (u8) v = (float)u8[(size_t)(y0 + r0) * W + x0 + 4 * j];
    const float4 w = make_float4(v, v + 1, v + 2, v + 3);
    float* pl[3] = {a, b, c};
#pragma unroll
    for (int k = 0; k < 3; ++k)
      for (int r = r0; r < TH; r += RPP)
        *reinterpret_cast<float4*>(pl[k] + (size_t)(y0 + r) * W + x0 + 4 * j) = w;
  }
}

int main() {
  const size_t px = 3840ull * 2160, n4 = px / 4;
  const int SETS = 3;
  float4* pl[SETS][3];
  uchar4* u8[SETS];
  for (int s = 0; s < SETS; ++s) {
    for (int k = 0; k < 3; ++k) CK(cudaMalloc(&pl[s][k], px * 4));
    CK(cudaMalloc(&u8[s], px));
    CK(cudaMemset(u8[s], 7 * s + 1, px));
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int REPS = 60;
  float ms;
  for (int grid_mult : {2, 4, 8, 16}) {
    const int grid = 148 * grid_mult, block = 256;
    for (int i = 0; i < 6; ++i) fill3<<<grid, block>>>(pl[i % SETS][0], pl[i % SETS][1], pl[i % SETS][2], n4, (float)i);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < REPS; ++i) fill3<<<grid, block>>>(pl[i % SETS][0], pl[i % SETS][1], pl[i % SETS][2], n4, (float)i);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("fill     grid %4d: %6.2f us / launch  %7.1f GB/s written\n", grid, ms / REPS * 1e3, 12.0 * px / (ms / REPS * 1e-3) / 1e9);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < REPS; ++i) fill3_u8<<<grid, block>>>(u8[i % SETS], pl[i % SETS][0], pl[i % SETS][1], pl[i % SETS][2], n4);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("fill+u8  grid %4d: %6.2f us / launch  %7.1f GB/s (13 B / px)\n", grid, ms / REPS * 1e3, 13.0 * px / (ms / REPS * 1e-3) / 1e9);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < REPS; ++i) {
      const int s = i % SETS, t = (i + 1) % SETS;
      copy4<<<grid, block>>>(pl[s][0], pl[t][1], n4);                     // one plane each way: 66 MB of traffic
    }
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("copy     grid %4d: %6.2f us / launch  %7.1f GB/s (read + write)\n", grid, ms / REPS * 1e3, 2.0 * n4 * 16 / (ms / REPS * 1e-3) / 1e9);
  }
  {
    unsigned* ctr; CK(cudaMalloc(&ctr, 4)); CK(cudaMemset(ctr, 0, 4));
    unsigned base = 0;
#define TILED(TW, TH, CPS, U8)                                                                              \
    {                                                                                                       \
      const int grid = 148 * CPS;                                                                           \
      for (int i = 0; i < 4; ++i) { fill3_tiled<TW, TH><<<grid, 256>>>((float*)pl[i % SETS][0], (float*)pl[i % SETS][1], (float*)pl[i % SETS][2], 3840, 2160, ctr, base, U8 ? (const unsigned char*)u8[i % SETS] : nullptr); base += (3840 / TW) * (2160 / TH) + grid; } \
      CK(cudaEventRecord(e0));                                                                              \
      for (int i = 0; i < REPS; ++i) { fill3_tiled<TW, TH><<<grid, 256>>>((float*)pl[i % SETS][0], (float*)pl[i % SETS][1], (float*)pl[i % SETS][2], 3840, 2160, ctr, base, U8 ? (const unsigned char*)u8[i % SETS] : nullptr); base += (3840 / TW) * (2160 / TH) + grid; } \
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));         \
      printf("tiled %3dx%-3d %d CTAs/SM%s: %6.2f us / launch  %7.1f GB/s written\n", TW, TH, CPS, U8 ? " +u8" : "    ", ms / REPS * 1e3, 12.0 * px / (ms / REPS * 1e-3) / 1e9); \
    }
    TILED(64, 48, 4, 0) TILED(64, 48, 4, 1) TILED(64, 16, 4, 0) TILED(128, 24, 4, 0) TILED(128, 48, 4, 0) TILED(256, 24, 4, 0) TILED(256, 48, 4, 0)
    TILED(64, 48, 8, 0) TILED(128, 24, 8, 0) TILED(256, 24, 8, 0) TILED(256, 24, 8, 1)
  }
  CK(cudaEventRecord(e0));
  for (int i = 0; i < REPS; ++i) for (int k = 0; k < 3; ++k) CK(cudaMemsetAsync(pl[i % SETS][k], i, px * 4));
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("memset x3         : %6.2f us / 3 planes %7.1f GB/s written\n", ms / REPS * 1e3, 12.0 * px / (ms / REPS * 1e-3) / 1e9);
  // one big copy as MEASURED_PEAKS does it (1 GiB elements would not fit the probe's budget: 512 MB each way)
  float4 *big0, *big1;
  const size_t nb4 = (512ull << 20) / 16;
  CK(cudaMalloc(&big0, nb4 * 16)); CK(cudaMalloc(&big1, nb4 * 16));
  CK(cudaMemset(big0, 1, nb4 * 16));
  for (int i = 0; i < 3; ++i) copy4<<<148 * 8, 256>>>(big0, big1, nb4);
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 10; ++i) copy4<<<148 * 8, 256>>>(big0, big1, nb4);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("copy 512 MB -> 512 MB: %7.1f GB/s (read + write)\n", 2.0 * nb4 * 16 / (ms / 10 * 1e-3) / 1e9);
  CK(cudaMemcpyAsync(big1, big0, nb4 * 16, cudaMemcpyDeviceToDevice));
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 10; ++i) CK(cudaMemcpyAsync(big1, big0, nb4 * 16, cudaMemcpyDeviceToDevice));
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("cudaMemcpy D2D 512 MB: %7.1f GB/s (read + write)\n", 2.0 * nb4 * 16 / (ms / 10 * 1e-3) / 1e9);
  return 0;
}
