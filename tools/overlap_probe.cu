// Does a kernel chain on one stream overlap with H2D band copies on another?  (pattern of the
// banded frame upload in klt_dev_build).  Timeline from %globaltimer written by the kernels
// themselves (no timing events); copies are stamped by a 1-thread kernel queued behind each.
#include <cstdio>
#include <chrono>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void spin(long long ns, unsigned long long* out) {
  const unsigned long long t0 = gtime();
  while (gtime() - t0 < (unsigned long long)ns) {}
  if (out && blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t0; out[1] = gtime(); }
}
__global__ void stamp(unsigned long long* out) { *out = gtime(); }
static double now() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv) {
  const int NB = argc > 1 ? atoi(argv[1]) : 4;
  const int grid = argc > 2 ? atoi(argv[2]) : 148;
  const int kus = argc > 3 ? atoi(argv[3]) : 5;
  const int mode = argc > 4 ? atoi(argv[4]) : 0;    // 0: kernels gated per band, 1: all copies first, 2: stamps via copy-stream kernels off
  cudaStream_t sc, sk; cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&sk, cudaStreamNonBlocking);
  unsigned char *h, *d; cudaMallocHost(&h, 8 << 20); cudaMalloc(&d, 8 << 20);
  unsigned long long *ts, *hts; cudaMalloc(&ts, 4096 * 8); cudaMallocHost(&hts, 4096 * 8);
  std::vector<cudaEvent_t> evb(NB);
  for (auto& e : evb) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  const size_t band = (8u << 20) / NB;
  for (int rep = 0; rep < 5; ++rep) {
    cudaDeviceSynchronize();
    cudaMemset(ts, 0, 4096 * 8);
    cudaDeviceSynchronize();
    const double t0 = now();
    stamp<<<1, 1, 0, sk>>>(ts);                       // origin
    for (int b = 0; b < NB; ++b) {
      cudaMemcpyAsync(d + b * band, h + b * band, band, cudaMemcpyHostToDevice, sc);
      if (mode != 2) stamp<<<1, 1, 0, sc>>>(ts + 1 + b);
      cudaEventRecord(evb[b], sc);
    }
    const double t1 = now();
    for (int b = 0; b < NB; ++b) {
      cudaStreamWaitEvent(sk, evb[mode == 1 ? NB - 1 : b], 0);
      for (int j = 0; j < 4; ++j) spin<<<grid, 256, 0, sk>>>(kus * 1000, ts + 64 + 2 * (4 * b + j));
    }
    const double t2 = now();
    cudaStreamSynchronize(sk);
    const double t3 = now();
    cudaMemcpy(hts, ts, 4096 * 8, cudaMemcpyDeviceToHost);
    printf("rep %d mode %d: copies enqueued in %.1f us, kernels enqueued by %.1f us, all done at %.1f us\n", rep, mode, t1 - t0, t2 - t0, t3 - t0);
    if (rep == 4) {
      for (int b = 0; b < NB; ++b) if (hts[1 + b]) printf("  copy %d done at %7.1f\n", b, (hts[1 + b] - hts[0]) * 1e-3);
      for (int i = 0; i < 4 * NB; ++i) printf("  kernel %2d %7.1f -> %7.1f\n", i, (hts[64 + 2 * i] - hts[0]) * 1e-3, (hts[65 + 2 * i] - hts[0]) * 1e-3);
    }
  }
  return 0;
}
