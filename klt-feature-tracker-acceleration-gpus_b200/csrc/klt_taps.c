/* klt_taps.c -- Gaussian / derivative-of-Gaussian taps on the host.
 *
 * Restates reference src/V1/convolve.c:60-130 (_computeKernels,
 * _KLTGetKernelWidths) and the sigma cache of convolve.c:24-27,287,310:
 * the reference keeps ONE tap pair in file statics and regenerates it only when
 * |sigma - sigma_last| > 0.05.  That rule is observable (two sigmas closer than
 * 0.05 share taps), so it is kept -- but per thread, so that contexts on
 * different host threads (one per GPU) do not interfere.
 */
#include <math.h>
#include <string.h>

#include "klt_internal.h"

#define MAXW KLT_DEV_MAX_TAPS

static __thread klt_dev_taps t_cache;
static __thread float t_sigma_last = -10.0f;

/* generate() is a pure function of sigma; the last few results are remembered so that the
 * per-frame sequence smooth -> pyramid -> gradient sigma does not redo ~200 exp() calls.
 * (This memo is invisible: the reference's cache RULE lives in klt_taps_for below.) */
#define MEMO 8
static __thread struct { float sigma; int used; klt_dev_taps taps; } t_memo[MEMO];
static __thread int t_memo_next = 0;

static void generate_uncached(float sigma, klt_dev_taps *t);

static void generate(float sigma, klt_dev_taps *t)
{
  int i;
  for (i = 0; i < MEMO; i++)
    if (t_memo[i].used && memcmp(&t_memo[i].sigma, &sigma, sizeof sigma) == 0) { *t = t_memo[i].taps; return; }
  generate_uncached(sigma, t);
  t_memo[t_memo_next].sigma = sigma;
  t_memo[t_memo_next].used = 1;
  t_memo[t_memo_next].taps = *t;
  t_memo_next = (t_memo_next + 1) % MEMO;
}

static void generate_uncached(float sigma, klt_dev_taps *t)
{
  const float cut = 0.01f;                 /* tails below 1 % of the peak are dropped */
  const int half = MAXW / 2;
  float g[MAXW], dg[MAXW];
  const float peak_g = 1.0f;
  const float peak_d = (float)(sigma * exp(-0.5f));
  int i, wg = MAXW, wd = MAXW;

  for (i = -half; i <= half; i++) {
    g[i + half] = (float)exp(-i * i / (2 * sigma * sigma));
    dg[i + half] = -i * g[i + half];
  }
  for (i = -half; fabs(g[i + half] / peak_g) < cut; i++) wg -= 2;
  for (i = -half; fabs(dg[i + half] / peak_d) < cut; i++) wd -= 2;
  if (wg == MAXW || wd == MAXW)
    KLTError("(_computeKernels) MAX_KERNEL_WIDTH %d is too small for "
             "a sigma of %f", MAXW, sigma);

  memset(t, 0, sizeof(*t));
  t->gauss_width = wg;
  t->deriv_width = wd;
  for (i = 0; i < wg; i++) t->gauss[i] = g[i + (MAXW - wg) / 2];
  for (i = 0; i < wd; i++) t->deriv[i] = dg[i + (MAXW - wd) / 2];

  {
    const int dh = wd / 2;
    float den = 0.0f;
    for (i = 0; i < wg; i++) den += t->gauss[i];
    for (i = 0; i < wg; i++) t->gauss[i] /= den;
    den = 0.0f;
    for (i = -dh; i <= dh; i++) den -= i * t->deriv[i + dh];
    for (i = -dh; i <= dh; i++) t->deriv[i + dh] /= den;
  }
}

void klt_taps_for(float sigma, klt_dev_taps *out)
{
  if (fabs(sigma - t_sigma_last) > 0.05) {
    generate(sigma, &t_cache);
    t_sigma_last = sigma;
  }
  *out = t_cache;
}

void _KLTGetKernelWidths(float sigma, int *gauss_width, int *gaussderiv_width)
{
  generate(sigma, &t_cache);               /* always regenerates (convolve.c:127) */
  t_sigma_last = sigma;
  *gauss_width = t_cache.gauss_width;
  *gaussderiv_width = t_cache.deriv_width;
}
