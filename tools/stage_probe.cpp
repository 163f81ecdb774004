// Host-side probe of the pageable-frame staging team (csrc/klt_stage.h): copy rate of a 4K frame
// into another buffer, chunk by chunk as klt_dev_build does, for 1..8 threads and several chunk
// sizes, with the idle gap between frames that a real call sequence has.
//   g++ -O2 -pthread -I klt-feature-tracker-acceleration-gpus_b200/csrc tools/stage_probe.cpp -o tools/stage_probe.bin
#include <cstdio>
#include <unistd.h>
#include "klt_stage.h"
int main() {
  const size_t N = 3840 * 2160;
  unsigned char* a = (unsigned char*)malloc(N);
  unsigned char* b = (unsigned char*)malloc(N);
  for (size_t i = 0; i < N; ++i) a[i] = (unsigned char)(i * 7 + 3);
  for (int gap_us : {0, 200, 2000})
    for (int T : {1, 2, 4, 8})
      for (size_t chunk : {(size_t)256 << 10, (size_t)1 << 20, (size_t)2 << 20, N}) {
        double best = 1e30, sum = 0;
        int ok = 1;
        for (int rep = 0; rep < 12; ++rep) {
          memset(b, rep, 4096);
          if (gap_us) usleep(gap_us);
          auto t0 = std::chrono::steady_clock::now();
          for (size_t o = 0; o < N; o += chunk) stage_team().copy(b + o, a + o, N - o < chunk ? N - o : chunk, T);
          auto t1 = std::chrono::steady_clock::now();
          const double us = std::chrono::duration<double, std::micro>(t1 - t0).count();
          if (rep >= 2) { sum += us; if (us < best) best = us; }
          ok = ok && memcmp(a, b, N) == 0;
        }
        printf("gap %4d us  threads %d  chunk %5zu KB: best %7.1f us  mean %7.1f us  (%.1f GB/s best) %s\n", gap_us, T,
               chunk >> 10, best, sum / 10, N / best * 1e-3, ok ? "" : "MISMATCH");
      }
  return 0;
}
