// CPU cost of CUDA submissions on this host: empty-kernel launches, event records, stream waits,
// pinned H2D enqueues, and one graph launch of a 20-kernel chain.
#include <cstdio>
#include <chrono>
#include <cuda_runtime.h>
__global__ void empty_kernel(int* p) { if (p && threadIdx.x == 12345) *p = 1; }
struct Big { float a[256]; };
__global__ void empty_big(Big b, int* p) { if (p && threadIdx.x == 12345) *p = (int)b.a[3]; }
static double now() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  cudaStream_t s, s2; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  cudaEvent_t ev; cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  const int N = 2000;
  for (int i = 0; i < 100; ++i) empty_kernel<<<1, 32, 0, s>>>(nullptr);
  cudaStreamSynchronize(s);
  double t0 = now();
  for (int i = 0; i < N; ++i) empty_kernel<<<1, 32, 0, s>>>(nullptr);
  double t1 = now(); cudaStreamSynchronize(s); double t2 = now();
  printf("empty kernel: enqueue %.2f us/launch, incl. drain %.2f us/launch\n", (t1 - t0) / N, (t2 - t0) / N);
  Big b; for (int i = 0; i < 256; ++i) b.a[i] = i;
  t0 = now();
  for (int i = 0; i < N; ++i) empty_big<<<1, 32, 0, s>>>(b, nullptr);
  t1 = now(); cudaStreamSynchronize(s); t2 = now();
  printf("1 KB-param kernel: enqueue %.2f us/launch, incl. drain %.2f us/launch\n", (t1 - t0) / N, (t2 - t0) / N);
  t0 = now();
  for (int i = 0; i < N; ++i) cudaEventRecord(ev, s);
  t1 = now(); cudaStreamSynchronize(s);
  printf("event record: %.2f us\n", (t1 - t0) / N);
  t0 = now();
  for (int i = 0; i < N; ++i) { cudaEventRecord(ev, s2); cudaStreamWaitEvent(s, ev, 0); }
  t1 = now(); cudaStreamSynchronize(s); cudaStreamSynchronize(s2);
  printf("record + cross-stream wait: %.2f us\n", (t1 - t0) / N);
  unsigned char *h, *d; cudaMallocHost(&h, 8 << 20); cudaMalloc(&d, 8 << 20);
  for (size_t sz : {4096ul, 49152ul, 1ul << 20, 8ul << 20}) {
    cudaStreamSynchronize(s);
    t0 = now();
    for (int i = 0; i < 200; ++i) cudaMemcpyAsync(d, h, sz, cudaMemcpyHostToDevice, s);
    t1 = now(); cudaStreamSynchronize(s); t2 = now();
    printf("H2D %8zu B: enqueue %.2f us, incl. drain %.2f us (%.1f GB/s)\n", sz, (t1 - t0) / 200, (t2 - t0) / 200, sz / ((t2 - t0) / 200) / 1e3);
  }
  // one sync round trip
  t0 = now();
  for (int i = 0; i < 200; ++i) { empty_kernel<<<1, 32, 0, s>>>(nullptr); cudaStreamSynchronize(s); }
  t1 = now();
  printf("launch + sync round trip: %.2f us\n", (t1 - t0) / 200);
  // graph of 20 kernels
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < 20; ++i) empty_big<<<1, 32, 0, s>>>(b, nullptr);
  cudaStreamEndCapture(s, &g);
  cudaGraphInstantiate(&ge, g, 0);
  for (int i = 0; i < 10; ++i) cudaGraphLaunch(ge, s);
  cudaStreamSynchronize(s);
  t0 = now();
  for (int i = 0; i < 200; ++i) cudaGraphLaunch(ge, s);
  t1 = now(); cudaStreamSynchronize(s); t2 = now();
  printf("graph of 20 kernels: enqueue %.2f us/graph, incl. drain %.2f us/graph (%.2f us/kernel)\n", (t1 - t0) / 200, (t2 - t0) / 200, (t2 - t0) / 200 / 20);
  t0 = now();
  for (int i = 0; i < 200; ++i) { cudaGraphLaunch(ge, s); cudaStreamSynchronize(s); }
  t1 = now();
  printf("graph launch + sync: %.2f us\n", (t1 - t0) / 200);
  return 0;
}
