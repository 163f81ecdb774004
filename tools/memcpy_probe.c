/* host memcpy bandwidth pageable -> pinned-like buffer, 1..8 threads (OpenMP), 8.3 MB frames */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <omp.h>
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e6 + t.tv_nsec * 1e-3; }
int main(void) {
  const size_t n = 3840 * 2160;
  unsigned char *src[8], *dst = (unsigned char *)malloc(n);
  for (int i = 0; i < 8; ++i) { src[i] = (unsigned char *)malloc(n); memset(src[i], i + 1, n); }
  memset(dst, 0, n);
  for (int nt = 1; nt <= 8; nt *= 2) {
    double best = 1e30;
    for (int rep = 0; rep < 24; ++rep) {
      const unsigned char *s = src[rep % 8];
      const double t0 = now();
#pragma omp parallel for num_threads(nt) schedule(static)
      for (int c = 0; c < 64; ++c) memcpy(dst + c * (n / 64), s + c * (n / 64), n / 64);
      const double t1 = now();
      if (rep >= 8 && t1 - t0 < best) best = t1 - t0;
    }
    printf("%d threads: %.1f us per 8.3 MB frame (%.1f GB/s)\n", nt, best, n / best / 1e3);
  }
  return 0;
}
