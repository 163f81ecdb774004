// klt_stage.h -- host-side staging team for pageable frames (included by klt_dev.cu and by
// tools/stage_probe.cpp).  Plain C++11, no CUDA.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

// The staged bytes are read next by the GPU's DMA engine, not by this CPU: streaming (non-temporal)
// stores put them straight into DRAM -- no read-for-ownership of the destination lines, nothing
// dirty left in this core's caches for the DMA reads to snoop out.  KLT_B200_STAGE_NT=0 keeps memcpy.
#if defined(__x86_64__)
__attribute__((target("avx2"))) static inline void klt_stream_copy_avx2(unsigned char* dst, const unsigned char* src, size_t n) {
  size_t head = (32 - ((uintptr_t)dst & 31)) & 31;
  if (head > n) head = n;
  if (head) { memcpy(dst, src, head); dst += head; src += head; n -= head; }
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256((const __m256i*)(src + i));
    const __m256i b = _mm256_loadu_si256((const __m256i*)(src + i + 32));
    const __m256i c = _mm256_loadu_si256((const __m256i*)(src + i + 64));
    const __m256i e = _mm256_loadu_si256((const __m256i*)(src + i + 96));
    _mm256_stream_si256((__m256i*)(dst + i), a);
    _mm256_stream_si256((__m256i*)(dst + i + 32), b);
    _mm256_stream_si256((__m256i*)(dst + i + 64), c);
    _mm256_stream_si256((__m256i*)(dst + i + 96), e);
  }
  for (; i + 32 <= n; i += 32) _mm256_stream_si256((__m256i*)(dst + i), _mm256_loadu_si256((const __m256i*)(src + i)));
  if (i < n) memcpy(dst + i, src + i, n - i);
  _mm_sfence();
}
#endif
static inline void klt_stage_copy(unsigned char* dst, const unsigned char* src, size_t n) {
#if defined(__x86_64__)
  static const int nt = (getenv("KLT_B200_STAGE_NT") ? atoi(getenv("KLT_B200_STAGE_NT")) : 1) && __builtin_cpu_supports("avx2");
  if (nt && n >= 4096) { klt_stream_copy_avx2(dst, src, n); return; }
#endif
  memcpy(dst, src, n);
}

// Staging team: a few process-wide helper threads that copy slices of a chunk next to the calling
// thread.  They spin on a generation counter for a short while after their last job (a frame is
// staged chunk by chunk, a few microseconds apart; consecutive frames ~200 us apart) and sleep on
// a condition variable otherwise, so a chunk costs one atomic round trip, not a thread wake-up.
struct StageTeam {
  static constexpr int MAXW = 15;
  std::mutex mu; std::condition_variable cv;
  std::vector<std::thread> workers;
  std::atomic<unsigned> gen{0};                 // bumped per job
  std::atomic<int> pending{0};                  // helper slices not finished yet
  std::atomic<int> sleepers{0};
  unsigned char* dst = nullptr; const unsigned char* src = nullptr; size_t bytes = 0; int parts = 1;
  bool stop = false;
  std::mutex job_mu;                            // one job at a time (contexts on several host threads share the team)
  static int spin_us() {                        // how long a helper spins for the next chunk before it sleeps
    static int us = getenv("KLT_B200_STAGE_SPIN_US") ? atoi(getenv("KLT_B200_STAGE_SPIN_US")) : 300;
    return us;
  }
  static void slice(unsigned char* dst, const unsigned char* src, size_t bytes, int parts, int k) {
    const size_t unit = (bytes / parts + 63) & ~(size_t)63;
    const size_t a = unit * k < bytes ? unit * k : bytes;
    const size_t b = k == parts - 1 ? bytes : (unit * (k + 1) < bytes ? unit * (k + 1) : bytes);
    if (b > a) klt_stage_copy(dst + a, src + a, b - a);
  }
  void run(int id) {
    unsigned seen = 0;
    for (;;) {
      unsigned g = gen.load(std::memory_order_acquire);
      if (g == seen) {
        const auto t0 = std::chrono::steady_clock::now();
        while ((g = gen.load(std::memory_order_acquire)) == seen) {
          if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(spin_us())) {
            std::unique_lock<std::mutex> lk(mu);
            sleepers.fetch_add(1);
            cv.wait(lk, [&] { return stop || gen.load(std::memory_order_acquire) != seen; });
            sleepers.fetch_sub(1);
            if (stop) return;
            g = gen.load(std::memory_order_acquire);
            break;
          }
          __builtin_ia32_pause();
        }
      }
      seen = g;
      if (id + 1 < parts) slice(dst, src, bytes, parts, id + 1);
      pending.fetch_sub(1, std::memory_order_acq_rel);
    }
  }
  void copy(unsigned char* d, const unsigned char* s, size_t n, int nthreads) {
    if (nthreads > MAXW + 1) nthreads = MAXW + 1;
    if (nthreads <= 1 || n < (256u << 10)) { klt_stage_copy(d, s, n); return; }
    std::lock_guard<std::mutex> job(job_mu);
    while ((int)workers.size() < nthreads - 1) { const int id = (int)workers.size(); workers.emplace_back([this, id] { run(id); }); }
    dst = d; src = s; bytes = n; parts = nthreads;
    pending.store((int)workers.size(), std::memory_order_release);     // every helper acknowledges the job
    gen.fetch_add(1, std::memory_order_acq_rel);
    if (sleepers.load() > 0) { std::lock_guard<std::mutex> lk(mu); cv.notify_all(); }
    slice(d, s, n, nthreads, 0);
    while (pending.load(std::memory_order_acquire) > 0) __builtin_ia32_pause();
  }
  ~StageTeam() {
    { std::lock_guard<std::mutex> lk(mu); stop = true; gen.fetch_add(1); }
    cv.notify_all();
    for (auto& t : workers) if (t.joinable()) t.detach();   // process exit: do not wait for sleepers
  }
};
static StageTeam& stage_team() { static StageTeam* t = new StageTeam(); return *t; }   // (leaked on purpose: no exit-order issues)
