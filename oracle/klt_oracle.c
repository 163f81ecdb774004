/* oracle/klt_oracle.c
 *
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  See klt_oracle.h.
 *
 * Plain-C restatement of the reference CPU path of the KLT tracker
 * (Birchfield's library as shipped in the reference repo, src/V1 == src/V3 CPU
 * sources).  Every function cites the reference file:line it follows.  The
 * arithmetic (operand types, promotion to double, summation order) is kept
 * identical so that this file is bit-exact against the compiled reference
 * (oracle/_ref) -- tests/test_oracle.py checks that, and checks the golden
 * feature table src/V1/feat/features2.ft byte for byte.
 *
 * Build: gcc -O2 -ffp-contract=off (no FMA, no fast-math).
 */
#include "klt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* taps -- reference src/V1/convolve.c:60-114 (_computeKernels)             */
/* ------------------------------------------------------------------------ */

typedef struct {
  int   wg, wd;
  float g[KLTO_MAX_TAPS];
  float d[KLTO_MAX_TAPS];
} tapset;

/* The reference keeps one tap set and the sigma it was made for in file
 * statics (convolve.c:24-27) and only regenerates when
 * fabs(sigma - sigma_last) > 0.05 (convolve.c:287, :310). */
static tapset s_taps;
static float  s_sigma_last = -10.0f;

static int make_taps(float sigma, tapset *t)
{
  const float factor = 0.01f;
  const int   hw = KLTO_MAX_TAPS / 2;
  float g[KLTO_MAX_TAPS], d[KLTO_MAX_TAPS];
  const float max_gauss = 1.0f;
  const float max_deriv = (float)(sigma * exp(-0.5f));
  int i, wg, wd;

  for (i = -hw; i <= hw; i++) {
    g[i + hw] = (float)exp(-i * i / (2 * sigma * sigma));
    d[i + hw] = -i * g[i + hw];
  }

  /* trim tails that are below 1 % of the peak, two taps at a time */
  wg = KLTO_MAX_TAPS;
  for (i = -hw; fabs(g[i + hw] / max_gauss) < factor; i++) wg -= 2;
  wd = KLTO_MAX_TAPS;
  for (i = -hw; fabs(d[i + hw] / max_deriv) < factor; i++) wd -= 2;
  if (wg == KLTO_MAX_TAPS || wd == KLTO_MAX_TAPS) return -1;

  for (i = 0; i < wg; i++) t->g[i] = g[i + (KLTO_MAX_TAPS - wg) / 2];
  for (i = 0; i < wd; i++) t->d[i] = d[i + (KLTO_MAX_TAPS - wd) / 2];
  t->wg = wg;
  t->wd = wd;

  {
    const int dhw = wd / 2;
    float den = 0.0f;
    for (i = 0; i < wg; i++) den += t->g[i];
    for (i = 0; i < wg; i++) t->g[i] /= den;
    den = 0.0f;
    for (i = -dhw; i <= dhw; i++) den -= i * t->d[i + dhw];
    for (i = -dhw; i <= dhw; i++) t->d[i + dhw] /= den;
  }
  return 0;
}

static void taps_if_needed(float sigma)
{
  if (fabs(sigma - s_sigma_last) > 0.05) {
    make_taps(sigma, &s_taps);
    s_sigma_last = sigma;
  }
}

/* reference convolve.c:122-130 (_KLTGetKernelWidths): always recomputes and
 * therefore also moves sigma_last. */
int klto_taps(float sigma, float *gauss, int *gauss_width,
              float *deriv, int *deriv_width)
{
  if (make_taps(sigma, &s_taps) != 0) return -1;
  s_sigma_last = sigma;
  if (gauss) memcpy(gauss, s_taps.g, sizeof(float) * s_taps.wg);
  if (deriv) memcpy(deriv, s_taps.d, sizeof(float) * s_taps.wd);
  if (gauss_width) *gauss_width = s_taps.wg;
  if (deriv_width) *deriv_width = s_taps.wd;
  return 0;
}

void klto_reset_tap_cache(void) { s_sigma_last = -10.0f; }

/* ------------------------------------------------------------------------ */
/* image stages                                                             */
/* ------------------------------------------------------------------------ */

/* reference convolve.c:37-53 (_KLTToFloatImage) */
void klto_to_float(const unsigned char *img, int ncols, int nrows, float *out)
{
  long i, n = (long)ncols * nrows;
  for (i = 0; i < n; i++) out[i] = (float)img[i];
}

/* reference convolve.c:137-182 (_convolveImageHoriz): true convolution, the
 * tap array is walked backwards while the pixels are walked forwards; columns
 * closer than the radius to either edge are written as 0. */
static void conv_rows(const float *in, int ncols, int nrows,
                      const float *k, int w, float *out)
{
  const int r = w / 2;
  int x, y, m;
  for (y = 0; y < nrows; y++) {
    const float *row = in + (long)y * ncols;
    float *o = out + (long)y * ncols;
    for (x = 0; x < ncols; x++) {
      if (x < r || x >= ncols - r) { o[x] = 0.0f; continue; }
      float sum = 0.0f;
      for (m = 0; m < w; m++) sum += row[x - r + m] * k[w - 1 - m];
      o[x] = sum;
    }
  }
}

/* reference convolve.c:189-242 (_convolveImageVert) */
static void conv_cols(const float *in, int ncols, int nrows,
                      const float *k, int w, float *out)
{
  const int r = w / 2;
  int x, y, m;
  for (y = 0; y < nrows; y++) {
    float *o = out + (long)y * ncols;
    if (y < r || y >= nrows - r) {
      for (x = 0; x < ncols; x++) o[x] = 0.0f;
      continue;
    }
    for (x = 0; x < ncols; x++) {
      float sum = 0.0f;
      for (m = 0; m < w; m++) sum += in[(long)(y - r + m) * ncols + x] * k[w - 1 - m];
      o[x] = sum;
    }
  }
}

/* reference convolve.c:249-266 (_convolveSeparate) */
void klto_convolve_separate(const float *in, int ncols, int nrows,
                            const float *kh, int wh, const float *kv, int wv,
                            float *out)
{
  float *tmp = (float *)malloc(sizeof(float) * (size_t)ncols * nrows);
  conv_rows(in, ncols, nrows, kh, wh, tmp);
  conv_cols(tmp, ncols, nrows, kv, wv, out);
  free(tmp);
}

/* reference convolve.c:300-314 (_KLTComputeSmoothedImage) */
void klto_smooth(const float *in, int ncols, int nrows, float sigma, float *out)
{
  taps_if_needed(sigma);
  klto_convolve_separate(in, ncols, nrows, s_taps.g, s_taps.wg, s_taps.g, s_taps.wg, out);
}

/* reference convolve.c:273-293 (_KLTComputeGradients) */
void klto_gradients(const float *in, int ncols, int nrows, float sigma,
                    float *gx, float *gy)
{
  taps_if_needed(sigma);
  klto_convolve_separate(in, ncols, nrows, s_taps.d, s_taps.wd, s_taps.g, s_taps.wg, gx);
  klto_convolve_separate(in, ncols, nrows, s_taps.g, s_taps.wg, s_taps.d, s_taps.wd, gy);
}

/* reference pyramid.c:87-131 (_KLTComputePyramid), one level step */
void klto_pyr_down(const float *in, int ncols, int nrows, int ss, float sigma,
                   float *out)
{
  const int oc = ncols / ss, orows = nrows / ss, half = ss / 2;
  float *tmp = (float *)malloc(sizeof(float) * (size_t)ncols * nrows);
  int x, y;
  klto_smooth(in, ncols, nrows, sigma, tmp);
  for (y = 0; y < orows; y++)
    for (x = 0; x < oc; x++)
      out[(long)y * oc + x] = tmp[(long)(ss * y + half) * ncols + (ss * x + half)];
  free(tmp);
}

/* ------------------------------------------------------------------------ */
/* parameters -- reference klt.c                                            */
/* ------------------------------------------------------------------------ */

static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

static void fix_window(klto_params *p)
{
  /* klt.c:298-317 / trackFeatures.c:1258-1278: odd, >= 3 */
  if (p->window_width % 2 != 1) p->window_width++;
  if (p->window_height % 2 != 1) p->window_height++;
  if (p->window_width < 3) p->window_width = 3;
  if (p->window_height < 3) p->window_height = 3;
}

/* reference klt_util.c:20-24 */
float klto_smooth_sigma(const klto_params *p)
{
  return p->smooth_sigma_fact * imax(p->window_width, p->window_height);
}

/* reference klt.c:288-343 */
void klto_change_pyramid(klto_params *p, int search_range)
{
  float window_halfwidth, subsampling;
  fix_window(p);
  window_halfwidth = imin(p->window_width, p->window_height) / 2.0f;
  subsampling = ((float)search_range) / window_halfwidth;
  if (subsampling < 1.0) {
    p->nPyramidLevels = 1;
  } else if (subsampling <= 3.0) {
    p->nPyramidLevels = 2; p->subsampling = 2;
  } else if (subsampling <= 5.0) {
    p->nPyramidLevels = 2; p->subsampling = 4;
  } else if (subsampling <= 9.0) {
    p->nPyramidLevels = 2; p->subsampling = 8;
  } else {
    float val = (float)(log(7.0 * subsampling + 1.0) / log(8.0));
    p->nPyramidLevels = (int)(val + 0.99);
    p->subsampling = 8;
  }
}

/* reference klt.c:362-431 */
void klto_update_border(klto_params *p)
{
  int wg, wd, smooth_hw, pyr_hw, n_invalid, window_hw, ss_power, i;
  const int ss = p->subsampling;
  fix_window(p);
  window_hw = imax(p->window_width, p->window_height) / 2;
  klto_taps(klto_smooth_sigma(p), NULL, &wg, NULL, &wd);
  smooth_hw = wg / 2;
  klto_taps(p->pyramid_sigma_fact * p->subsampling, NULL, &wg, NULL, &wd);
  pyr_hw = wg / 2;
  n_invalid = smooth_hw;
  for (i = 1; i < p->nPyramidLevels; i++) {
    float val = ((float)n_invalid + pyr_hw) / ss;
    n_invalid = (int)(val + 0.99);
  }
  ss_power = 1;
  for (i = 1; i < p->nPyramidLevels; i++) ss_power *= ss;
  p->borderx = p->bordery = (n_invalid + window_hw) * ss_power;
}

/* reference klt.c:20-44, :90-135 */
void klto_default_params(klto_params *p)
{
  p->mindist = 10;
  p->window_width = p->window_height = 7;
  p->smoothBeforeSelecting = 1;
  p->min_eigenvalue = 1;
  p->min_determinant = 0.01f;
  p->min_displacement = 0.1f;
  p->max_iterations = 10;
  p->max_residue = 10.0f;
  p->grad_sigma = 1.0f;
  p->smooth_sigma_fact = 0.1f;
  p->pyramid_sigma_fact = 0.9f;
  p->step_factor = 1.0f;
  p->nSkippedPixels = 0;
  p->nPyramidLevels = 0;
  p->subsampling = 0;
  p->lighting_insensitive = 0;
  klto_change_pyramid(p, 15);
  klto_update_border(p);
}

/* ------------------------------------------------------------------------ */
/* pyramids of one frame -- reference trackFeatures.c:1309-1321             */
/* ------------------------------------------------------------------------ */

struct klto_pyramids {
  int    nlevels;
  int    ncols[32], nrows[32];
  float *img[32], *gx[32], *gy[32];
};

klto_pyramids *klto_build_pyramids(const unsigned char *img, int ncols, int nrows,
                                   const klto_params *p)
{
  klto_pyramids *q = (klto_pyramids *)calloc(1, sizeof(*q));
  const int ss = p->subsampling;
  const float pyr_sigma = ss * p->pyramid_sigma_fact;
  size_t n = (size_t)ncols * nrows;
  float *tmp = (float *)malloc(sizeof(float) * n);
  int l, w = ncols, h = nrows;

  q->nlevels = p->nPyramidLevels;
  klto_to_float(img, ncols, nrows, tmp);
  for (l = 0; l < q->nlevels; l++) {
    q->ncols[l] = w; q->nrows[l] = h;
    q->img[l] = (float *)malloc(sizeof(float) * (size_t)w * h);
    q->gx[l]  = (float *)malloc(sizeof(float) * (size_t)w * h);
    q->gy[l]  = (float *)malloc(sizeof(float) * (size_t)w * h);
    w /= ss; h /= ss;
  }
  klto_smooth(tmp, ncols, nrows, klto_smooth_sigma(p), q->img[0]);
  for (l = 1; l < q->nlevels; l++)
    klto_pyr_down(q->img[l - 1], q->ncols[l - 1], q->nrows[l - 1], ss, pyr_sigma, q->img[l]);
  for (l = 0; l < q->nlevels; l++)
    klto_gradients(q->img[l], q->ncols[l], q->nrows[l], p->grad_sigma, q->gx[l], q->gy[l]);
  free(tmp);
  return q;
}

void klto_free_pyramids(klto_pyramids *q)
{
  int l;
  if (!q) return;
  for (l = 0; l < q->nlevels; l++) { free(q->img[l]); free(q->gx[l]); free(q->gy[l]); }
  free(q);
}

int klto_pyr_levels(const klto_pyramids *q) { return q->nlevels; }
void klto_pyr_dims(const klto_pyramids *q, int level, int *ncols, int *nrows)
{
  *ncols = q->ncols[level]; *nrows = q->nrows[level];
}
const float *klto_pyr_data(const klto_pyramids *q, int which, int level)
{
  return which == 0 ? q->img[level] : which == 1 ? q->gx[level] : q->gy[level];
}

/* ------------------------------------------------------------------------ */
/* tracker -- reference trackFeatures.c                                     */
/* ------------------------------------------------------------------------ */

/* reference trackFeatures.c:31-57 (_interpolate) */
static float bilinear(float x, float y, const float *img, int ncols)
{
  const int xt = (int)x, yt = (int)y;
  const float ax = x - xt, ay = y - yt;
  const float *p = img + (long)ncols * yt + xt;
  return ((1 - ax) * (1 - ay) * p[0] +
          ax * (1 - ay) * p[1] +
          (1 - ax) * ay * p[ncols] +
          ax * ay * p[ncols + 1]);
}

static int window_oob(float x, float y, int hw, int hh, int nc, int nr)
{
  const float one_plus_eps = 1.001f;
  return (x - hw < 0.0f || nc - (x + hw) < one_plus_eps ||
          y - hh < 0.0f || nr - (y + hh) < one_plus_eps);
}

/* reference trackFeatures.c:132-167 (_computeIntensityDifferenceLightingInsensitive):
 * gain alpha = sqrt(mean(g1^2) / mean(g2^2)), bias belta = mean(g1) - alpha * mean(g2) over the
 * window, imgdiff = g1 - g2 * alpha - belta.  All sums are sequential float sums in raster order;
 * the sqrt is the double sqrt of a float quotient. */
static void intensity_difference_li(const float *img1, int nc1, const float *img2, int nc2, float x1, float y1,
                                    float x2, float y2, int ww, int wh, float *diff)
{
  const int hw = ww / 2, hh = wh / 2;
  float g1, g2, sum1_squared = 0, sum2_squared = 0, sum1 = 0, sum2 = 0;
  float mean1, mean2, alpha, belta;
  int i, j;
  for (j = -hh; j <= hh; j++)
    for (i = -hw; i <= hw; i++) {
      g1 = bilinear(x1 + i, y1 + j, img1, nc1);
      g2 = bilinear(x2 + i, y2 + j, img2, nc2);
      sum1 += g1; sum2 += g2;
      sum1_squared += g1 * g1;
      sum2_squared += g2 * g2;
    }
  mean1 = sum1_squared / (ww * wh);
  mean2 = sum2_squared / (ww * wh);
  alpha = (float)sqrt(mean1 / mean2);
  mean1 = sum1 / (ww * wh);
  mean2 = sum2 / (ww * wh);
  belta = mean1 - alpha * mean2;
  for (j = -hh; j <= hh; j++)
    for (i = -hw; i <= hw; i++) {
      g1 = bilinear(x1 + i, y1 + j, img1, nc1);
      g2 = bilinear(x2 + i, y2 + j, img2, nc2);
      *diff++ = g1 - g2 * alpha - belta;
    }
}

/* reference trackFeatures.c:178-220 (_computeGradientSumLightingInsensitive).  Its gain is the
 * square root of the ratio of the window MEANS (the variables are called sum*_squared but
 * accumulate g, not g*g, :202) -- restated as it is. */
static void gradient_sum_li(const float *gx1, const float *gy1, const float *gx2, const float *gy2,
                            const float *img1, int nc1, const float *img2, int nc2, float x1, float y1,
                            float x2, float y2, int ww, int wh, float *wx, float *wy)
{
  const int hw = ww / 2, hh = wh / 2;
  float g1, g2, sum1_squared = 0, sum2_squared = 0, mean1, mean2, alpha;
  int i, j;
  for (j = -hh; j <= hh; j++)
    for (i = -hw; i <= hw; i++) {
      g1 = bilinear(x1 + i, y1 + j, img1, nc1);
      g2 = bilinear(x2 + i, y2 + j, img2, nc2);
      sum1_squared += g1; sum2_squared += g2;
    }
  mean1 = sum1_squared / (ww * wh);
  mean2 = sum2_squared / (ww * wh);
  alpha = (float)sqrt(mean1 / mean2);
  for (j = -hh; j <= hh; j++)
    for (i = -hw; i <= hw; i++) {
      g1 = bilinear(x1 + i, y1 + j, gx1, nc1);
      g2 = bilinear(x2 + i, y2 + j, gx2, nc2);
      *wx++ = g1 + g2 * alpha;
      g1 = bilinear(x1 + i, y1 + j, gy1, nc1);
      g2 = bilinear(x2 + i, y2 + j, gy2, nc2);
      *wy++ = g1 + g2 * alpha;
    }
}

/* reference trackFeatures.c:381-486 (_trackFeature), translation model; lighting != 0 takes the
 * gain / bias normalised windows (:433-437, :466-468). */
int klto_track_level_li(float x1, float y1, float *x2, float *y2,
                        const float *img1, const float *gx1, const float *gy1,
                        const float *img2, const float *gx2, const float *gy2,
                        int nc, int nr, int ww, int wh, float step_factor,
                        int max_iterations, float small, float th, float max_residue, int lighting);

int klto_track_level(float x1, float y1, float *x2, float *y2,
                     const float *img1, const float *gx1, const float *gy1,
                     const float *img2, const float *gx2, const float *gy2,
                     int nc, int nr, int ww, int wh, float step_factor,
                     int max_iterations, float small, float th, float max_residue)
{
  return klto_track_level_li(x1, y1, x2, y2, img1, gx1, gy1, img2, gx2, gy2, nc, nr, ww, wh, step_factor,
                             max_iterations, small, th, max_residue, 0);
}

int klto_track_level_li(float x1, float y1, float *x2, float *y2,
                        const float *img1, const float *gx1, const float *gy1,
                        const float *img2, const float *gx2, const float *gy2,
                        int nc, int nr, int ww, int wh, float step_factor,
                        int max_iterations, float small, float th, float max_residue, int lighting)
{
  const int hw = ww / 2, hh = wh / 2, npix = ww * wh;
  float *diff = (float *)malloc(sizeof(float) * npix);
  float *wx = (float *)malloc(sizeof(float) * npix);
  float *wy = (float *)malloc(sizeof(float) * npix);
  float gxx, gxy, gyy, ex, ey, dx = 0.0f, dy = 0.0f;
  int iteration = 0, status, i, j, k;

  do {
    /* :418-425 both windows must lie inside the level image */
    if (window_oob(x1, y1, hw, hh, nc, nr) || window_oob(*x2, *y2, hw, hh, nc, nr)) {
      status = KLTO_OOB;
      break;
    }
    /* :68-87 and :98-123 intensity difference and gradient sum windows */
    k = 0;
    if (lighting) {
      intensity_difference_li(img1, nc, img2, nc, x1, y1, *x2, *y2, ww, wh, diff);
      gradient_sum_li(gx1, gy1, gx2, gy2, img1, nc, img2, nc, x1, y1, *x2, *y2, ww, wh, wx, wy);
    } else
    for (j = -hh; j <= hh; j++)
      for (i = -hw; i <= hw; i++, k++) {
        float a = bilinear(x1 + i, y1 + j, img1, nc);
        float b = bilinear(*x2 + i, *y2 + j, img2, nc);
        diff[k] = a - b;
        a = bilinear(x1 + i, y1 + j, gx1, nc);
        b = bilinear(*x2 + i, *y2 + j, gx2, nc);
        wx[k] = a + b;
        a = bilinear(x1 + i, y1 + j, gy1, nc);
        b = bilinear(*x2 + i, *y2 + j, gy2, nc);
        wy[k] = a + b;
      }
    /* :227-249 */
    gxx = 0.0f; gxy = 0.0f; gyy = 0.0f;
    for (k = 0; k < npix; k++) {
      gxx += wx[k] * wx[k];
      gxy += wx[k] * wy[k];
      gyy += wy[k] * wy[k];
    }
    /* :257-279 */
    ex = 0.0f; ey = 0.0f;
    for (k = 0; k < npix; k++) {
      ex += diff[k] * wx[k];
      ey += diff[k] * wy[k];
    }
    ex *= step_factor;
    ey *= step_factor;
    /* :293-307 */
    {
      const float det = gxx * gyy - gxy * gxy;
      if (det < small) { status = KLTO_SMALL_DET; break; }
      dx = (gyy * ex - gxy * ey) / det;
      dy = (gxx * ey - gxy * ex) / det;
      status = KLTO_TRACKED;
    }
    *x2 += dx;
    *y2 += dy;
    iteration++;
  } while ((fabs(dx) >= th || fabs(dy) >= th) && iteration < max_iterations);

  /* :459-462 */
  if (window_oob(*x2, *y2, hw, hh, nc, nr)) status = KLTO_OOB;

  /* :464-474 residue */
  if (status == KLTO_TRACKED) {
    float sum = 0.0f;
    k = 0;
    if (lighting)
      intensity_difference_li(img1, nc, img2, nc, x1, y1, *x2, *y2, ww, wh, diff);
    else
    for (j = -hh; j <= hh; j++)
      for (i = -hw; i <= hw; i++, k++)
        diff[k] = bilinear(x1 + i, y1 + j, img1, nc) - bilinear(*x2 + i, *y2 + j, img2, nc);
    for (k = 0; k < npix; k++) sum += (float)fabs(diff[k]);
    if (sum / (ww * wh) > max_residue) status = KLTO_LARGE_RESIDUE;
  }

  free(diff); free(wx); free(wy);

  /* :479-484 */
  if (status == KLTO_SMALL_DET) return KLTO_SMALL_DET;
  if (status == KLTO_OOB) return KLTO_OOB;
  if (status == KLTO_LARGE_RESIDUE) return KLTO_LARGE_RESIDUE;
  if (iteration >= max_iterations) return KLTO_MAX_ITERATIONS;
  return KLTO_TRACKED;
}

/* reference trackFeatures.c:1343-1437 (feature loop of KLTTrackFeatures) */
void klto_track(const klto_pyramids *p1, const klto_pyramids *p2,
                const klto_params *p, int n, float *x, float *y, int *val)
{
  const float ss = (float)p->subsampling;
  const int L = p->nPyramidLevels;
  const int ncols = p1->ncols[0], nrows = p1->nrows[0];
  int f, r;

  for (f = 0; f < n; f++) {
    float xloc, yloc, xout, yout;
    int v = KLTO_TRACKED;
    if (val[f] < 0) continue;
    xloc = x[f]; yloc = y[f];
    for (r = L - 1; r >= 0; r--) { xloc /= ss; yloc /= ss; }
    xout = xloc; yout = yloc;
    for (r = L - 1; r >= 0; r--) {
      xloc *= ss; yloc *= ss; xout *= ss; yout *= ss;
      v = klto_track_level_li(xloc, yloc, &xout, &yout,
                              p1->img[r], p1->gx[r], p1->gy[r],
                              p2->img[r], p2->gx[r], p2->gy[r],
                              p1->ncols[r], p1->nrows[r],
                              p->window_width, p->window_height, p->step_factor,
                              p->max_iterations, p->min_determinant,
                              p->min_displacement, p->max_residue, p->lighting_insensitive);
      if (v == KLTO_SMALL_DET || v == KLTO_OOB) break;
    }
    if (v == KLTO_OOB ||
        xout < p->borderx || xout > ncols - 1 - p->borderx ||
        yout < p->bordery || yout > nrows - 1 - p->bordery) {
      x[f] = -1.0f; y[f] = -1.0f; val[f] = KLTO_OOB;
    } else if (v == KLTO_SMALL_DET || v == KLTO_LARGE_RESIDUE || v == KLTO_MAX_ITERATIONS) {
      x[f] = -1.0f; y[f] = -1.0f; val[f] = v;
    } else {
      x[f] = xout; y[f] = yout; val[f] = KLTO_TRACKED;
    }
  }
}

/* ------------------------------------------------------------------------ */
/* affine consistency check -- reference trackFeatures.c:506-1224, :1438-1497 */
/* ------------------------------------------------------------------------ */

/* reference trackFeatures.c:546-604 (_am_gauss_jordan_elimination, n x n system, one right-hand
 * side): full pivoting, the pivot search compares |a| in double as fabs() does */
static int gauss_jordan(float a[6][6], int n, float b[6])
{
  int indxc[6], indxr[6], ipiv[6];
  int i, j, k, l, ll, col = 0, row = 0;
  float big, dum, pivinv, temp;
  for (j = 0; j < n; j++) ipiv[j] = 0;
  for (i = 0; i < n; i++) {
    big = 0.0f;
    for (j = 0; j < n; j++)
      if (ipiv[j] != 1)
        for (k = 0; k < n; k++) {
          if (ipiv[k] == 0) {
            if (fabs(a[j][k]) >= big) { big = (float)fabs(a[j][k]); row = j; col = k; }
          } else if (ipiv[k] > 1) return KLTO_SMALL_DET;
        }
    ++(ipiv[col]);
    if (row != col) {
      for (l = 0; l < n; l++) { temp = a[row][l]; a[row][l] = a[col][l]; a[col][l] = temp; }
      temp = b[row]; b[row] = b[col]; b[col] = temp;
    }
    indxr[i] = row; indxc[i] = col;
    if (a[col][col] == 0.0) return KLTO_SMALL_DET;
    pivinv = 1.0f / a[col][col];
    a[col][col] = 1.0f;
    for (l = 0; l < n; l++) a[col][l] *= pivinv;
    b[col] *= pivinv;
    for (ll = 0; ll < n; ll++)
      if (ll != col) {
        dum = a[ll][col];
        a[ll][col] = 0.0f;
        for (l = 0; l < n; l++) a[ll][l] -= a[col][l] * dum;
        b[ll] -= b[col] * dum;
      }
  }
  (void)indxr; (void)indxc;        /* the column unscramble (:595-599) only touches the inverse */
  return KLTO_TRACKED;
}

/* reference trackFeatures.c:952-1224 (_am_trackFeatureAffine).  img1/gx1/gy1: the feature's saved
 * (aw+2) x (ah+2) template; img2/gx2/gy2: level 0 of the new frame.  The result position is not
 * written back by the caller (:1490-1491), only the status and the mapping A. */
int klto_track_affine_feature(float x1, float y1, float *x2, float *y2,
                              const float *img1, const float *gx1, const float *gy1, int nc1, int nr1,
                              const float *img2, const float *gx2, const float *gy2, int nc2, int nr2,
                              int width, int height, float step_factor, int max_iterations,
                              float small, float th, float th_aff, float max_residue,
                              int affine_map, float mdd,
                              float *Axx, float *Ayx, float *Axy, float *Ayy, int lighting)
{
  const int hw = width / 2, hh = height / 2, npix = width * height;
  float *diff = (float *)malloc(sizeof(float) * npix);
  float *wx = (float *)malloc(sizeof(float) * npix);
  float *wy = (float *)malloc(sizeof(float) * npix);
  float gxx, gxy, gyy, ex, ey, dx = 0.0f, dy = 0.0f;
  const float one_plus_eps = 1.001f;
  const float old_x2 = *x2, old_y2 = *y2;
  int iteration = 0, status = 0, convergence = 0, i, j, k;
  float T[6][6], a[6];

  do {
    if (!affine_map) {
      /* :1010-1059 pure translation against the template */
      if (x1 - hw < 0.0f || nc1 - (x1 + hw) < one_plus_eps ||
          *x2 - hw < 0.0f || nc2 - (*x2 + hw) < one_plus_eps ||
          y1 - hh < 0.0f || nr1 - (y1 + hh) < one_plus_eps ||
          *y2 - hh < 0.0f || nr2 - (*y2 + hh) < one_plus_eps) { status = KLTO_OOB; break; }
      k = 0;
      if (lighting) {             /* :1024-1028 gain / bias normalised windows, template against frame */
        intensity_difference_li(img1, nc1, img2, nc2, x1, y1, *x2, *y2, width, height, diff);
        gradient_sum_li(gx1, gy1, gx2, gy2, img1, nc1, img2, nc2, x1, y1, *x2, *y2, width, height, wx, wy);
      } else
      for (j = -hh; j <= hh; j++)
        for (i = -hw; i <= hw; i++, k++) {
          diff[k] = bilinear(x1 + i, y1 + j, img1, nc1) - bilinear(*x2 + i, *y2 + j, img2, nc2);
          wx[k] = bilinear(x1 + i, y1 + j, gx1, nc1) + bilinear(*x2 + i, *y2 + j, gx2, nc2);
          wy[k] = bilinear(x1 + i, y1 + j, gy1, nc1) + bilinear(*x2 + i, *y2 + j, gy2, nc2);
        }
      gxx = 0.0f; gxy = 0.0f; gyy = 0.0f;
      for (k = 0; k < npix; k++) { gxx += wx[k] * wx[k]; gxy += wx[k] * wy[k]; gyy += wy[k] * wy[k]; }
      ex = 0.0f; ey = 0.0f;
      for (k = 0; k < npix; k++) { ex += diff[k] * wx[k]; ey += diff[k] * wy[k]; }
      ex *= step_factor; ey *= step_factor;
      {
        const float det = gxx * gyy - gxy * gxy;
        if (det < small) status = KLTO_SMALL_DET;
        else { dx = (gyy * ex - gxy * ey) / det; dy = (gxx * ey - gxy * ex) / det; status = KLTO_TRACKED; }
      }
      convergence = (fabs(dx) < th && fabs(dy) < th);
      *x2 += dx; *y2 += dy;
    } else {
      /* :1061-1186 affine tracker */
      float ul_x = *Axx * (-hw) + *Axy * hh + *x2, ul_y = *Ayx * (-hw) + *Ayy * hh + *y2;
      float ll_x = *Axx * (-hw) + *Axy * (-hh) + *x2, ll_y = *Ayx * (-hw) + *Ayy * (-hh) + *y2;
      float ur_x = *Axx * hw + *Axy * hh + *x2, ur_y = *Ayx * hw + *Ayy * hh + *y2;
      float lr_x = *Axx * hw + *Axy * (-hh) + *x2, lr_y = *Ayx * hw + *Ayy * (-hh) + *y2;
      if (x1 - hw < 0.0f || nc1 - (x1 + hw) < one_plus_eps ||
          y1 - hh < 0.0f || nr1 - (y1 + hh) < one_plus_eps ||
          ul_x < 0.0f || nc2 - ul_x < one_plus_eps || ll_x < 0.0f || nc2 - ll_x < one_plus_eps ||
          ur_x < 0.0f || nc2 - ur_x < one_plus_eps || lr_x < 0.0f || nc2 - lr_x < one_plus_eps ||
          ul_y < 0.0f || nr2 - ul_y < one_plus_eps || ll_y < 0.0f || nr2 - ll_y < one_plus_eps ||
          ur_y < 0.0f || nr2 - ur_y < one_plus_eps || lr_y < 0.0f || nr2 - lr_y < one_plus_eps) {
        status = KLTO_OOB; break;
      }
      /* :700-722 and :610-632: difference against the mapped window, gradients of frame 2 only */
      k = 0;
      for (j = -hh; j <= hh; j++)
        for (i = -hw; i <= hw; i++, k++) {
          const float g1 = bilinear(x1 + i, y1 + j, img1, nc1);
          const float mi = *Axx * i + *Axy * j, mj = *Ayx * i + *Ayy * j;
          diff[k] = g1 - bilinear(*x2 + mi, *y2 + mj, img2, nc2);
          wx[k] = bilinear(*x2 + mi, *y2 + mj, gx2, nc2);
          wy[k] = bilinear(*x2 + mi, *y2 + mj, gy2, nc2);
        }
      if (affine_map == 1) {
        /* :900-928 and :846-892 */
        for (i = 0; i < 4; i++) { a[i] = 0.0f; for (j = 0; j < 4; j++) T[i][j] = 0.0f; }
        k = 0;
        for (j = -hh; j <= hh; j++)
          for (i = -hw; i <= hw; i++, k++) {
            const float d = diff[k], dgx = d * wx[k], dgy = d * wy[k];
            a[0] += dgx * i + dgy * j;
            a[1] += dgy * i - dgx * j;
            a[2] += dgx;
            a[3] += dgy;
          }
        for (i = 0; i < 4; i++) a[i] *= 0.5;
        k = 0;
        for (j = -hh; j <= hh; j++)
          for (i = -hw; i <= hw; i++, k++) {
            const float gx = wx[k], gy = wy[k], x = (float)i, y = (float)j;
            T[0][0] += (x * gx + y * gy) * (x * gx + y * gy);
            T[0][1] += (x * gx + y * gy) * (x * gy - y * gx);
            T[0][2] += (x * gx + y * gy) * gx;
            T[0][3] += (x * gx + y * gy) * gy;
            T[1][1] += (x * gy - y * gx) * (x * gy - y * gx);
            T[1][2] += (x * gy - y * gx) * gx;
            T[1][3] += (x * gy - y * gx) * gy;
            T[2][2] += gx * gx;
            T[2][3] += gx * gy;
            T[3][3] += gy * gy;
          }
        for (j = 0; j < 3; j++) for (i = j + 1; i < 4; i++) T[i][j] = T[j][i];
        status = gauss_jordan(T, 4, a);
        *Axx += a[0]; *Ayx += a[1]; *Ayy = *Axx; *Axy = -(*Ayx);
        dx = a[2]; dy = a[3];
      } else {
        /* :806-838 and :730-798 */
        for (i = 0; i < 6; i++) { a[i] = 0.0f; for (j = 0; j < 6; j++) T[i][j] = 0.0f; }
        k = 0;
        for (j = -hh; j <= hh; j++)
          for (i = -hw; i <= hw; i++, k++) {
            const float d = diff[k], dgx = d * wx[k], dgy = d * wy[k];
            a[0] += dgx * i; a[1] += dgy * i; a[2] += dgx * j; a[3] += dgy * j; a[4] += dgx; a[5] += dgy;
          }
        for (i = 0; i < 6; i++) a[i] *= 0.5;
        k = 0;
        for (j = -hh; j <= hh; j++)
          for (i = -hw; i <= hw; i++, k++) {
            const float gx = wx[k], gy = wy[k];
            const float Gxx = gx * gx, Gxy = gx * gy, Gyy = gy * gy;
            const float x = (float)i, y = (float)j, xx = x * x, xy = x * y, yy = y * y;
            T[0][0] += xx * Gxx; T[0][1] += xx * Gxy; T[0][2] += xy * Gxx; T[0][3] += xy * Gxy;
            T[0][4] += x * Gxx;  T[0][5] += x * Gxy;
            T[1][1] += xx * Gyy; T[1][2] += xy * Gxy; T[1][3] += xy * Gyy; T[1][4] += x * Gxy; T[1][5] += x * Gyy;
            T[2][2] += yy * Gxx; T[2][3] += yy * Gxy; T[2][4] += y * Gxx;  T[2][5] += y * Gxy;
            T[3][3] += yy * Gyy; T[3][4] += y * Gxy;  T[3][5] += y * Gyy;
            T[4][4] += Gxx; T[4][5] += Gxy; T[5][5] += Gyy;
          }
        for (j = 0; j < 5; j++) for (i = j + 1; i < 6; i++) T[i][j] = T[j][i];
        status = gauss_jordan(T, 6, a);
        *Axx += a[0]; *Ayx += a[1]; *Axy += a[2]; *Ayy += a[3];
        dx = a[4]; dy = a[5];
      }
      *x2 += dx; *y2 += dy;
      /* :1162-1173 corner motion */
      ul_x -= *Axx * (-hw) + *Axy * hh + *x2;    ul_y -= *Ayx * (-hw) + *Ayy * hh + *y2;
      ll_x -= *Axx * (-hw) + *Axy * (-hh) + *x2; ll_y -= *Ayx * (-hw) + *Ayy * (-hh) + *y2;
      ur_x -= *Axx * hw + *Axy * hh + *x2;       ur_y -= *Ayx * hw + *Ayy * hh + *y2;
      lr_x -= *Axx * hw + *Axy * (-hh) + *x2;    lr_y -= *Ayx * hw + *Ayy * (-hh) + *y2;
      convergence = (fabs(dx) < th && fabs(dy) < th &&
                     fabs(ul_x) < th_aff && fabs(ul_y) < th_aff && fabs(ll_x) < th_aff && fabs(ll_y) < th_aff &&
                     fabs(ur_x) < th_aff && fabs(ur_y) < th_aff && fabs(lr_x) < th_aff && fabs(lr_y) < th_aff);
    }
    if (status == KLTO_SMALL_DET) break;
    iteration++;
  } while (!convergence && iteration < max_iterations);

  /* :1193-1200 */
  if (*x2 - hw < 0.0f || nc2 - (*x2 + hw) < one_plus_eps ||
      *y2 - hh < 0.0f || nr2 - (*y2 + hh) < one_plus_eps) status = KLTO_OOB;
  if ((*x2 - old_x2) > mdd || (*y2 - old_y2) > mdd) status = KLTO_OOB;
  /* :1203-1214 residue */
  if (status == KLTO_TRACKED) {
    float sum = 0.0f;
    k = 0;
    for (j = -hh; j <= hh; j++)
      for (i = -hw; i <= hw; i++, k++) {
        const float g1 = bilinear(x1 + i, y1 + j, img1, nc1);
        if (!affine_map) diff[k] = g1 - bilinear(*x2 + i, *y2 + j, img2, nc2);
        else {
          const float mi = *Axx * i + *Axy * j, mj = *Ayx * i + *Ayy * j;
          diff[k] = g1 - bilinear(*x2 + mi, *y2 + mj, img2, nc2);
        }
      }
    for (k = 0; k < npix; k++) sum += (float)fabs(diff[k]);
    if (sum / (width * height) > max_residue) status = KLTO_LARGE_RESIDUE;
  }
  free(diff); free(wx); free(wy);
  return status;
}

/* reference trackFeatures.c:1343-1497: the feature loop of KLTTrackFeatures with
 * tc->affineConsistencyCheck >= 0.  st[f] and tmpl (n x 3 x (aw+2)(ah+2) floats: image, gradx,
 * grady of the saved template) carry the per-feature fields aff_img*, aff_x .. aff_Ayy. */
void klto_track_affine(const klto_pyramids *p1, const klto_pyramids *p2, const klto_params *p,
                       const klto_affine_params *ap, int n, float *x, float *y, int *val,
                       klto_affine_state *st, float *tmpl)
{
  const float ss = (float)p->subsampling;
  const int L = p->nPyramidLevels;
  const int ncols = p1->ncols[0], nrows = p1->nrows[0];
  const int tw = ap->window_width + 2, th = ap->window_height + 2, tsz = tw * th;
  int f, r, i, j;

  for (f = 0; f < n; f++) {
    float xloc, yloc, xout, yout;
    int v = KLTO_TRACKED;
    float *t_img = tmpl + (size_t)f * 3 * tsz, *t_gx = t_img + tsz, *t_gy = t_gx + tsz;
    if (val[f] < 0) continue;
    xloc = x[f]; yloc = y[f];
    for (r = L - 1; r >= 0; r--) { xloc /= ss; yloc /= ss; }
    xout = xloc; yout = yloc;
    for (r = L - 1; r >= 0; r--) {
      xloc *= ss; yloc *= ss; xout *= ss; yout *= ss;
      v = klto_track_level_li(xloc, yloc, &xout, &yout, p1->img[r], p1->gx[r], p1->gy[r],
                              p2->img[r], p2->gx[r], p2->gy[r], p1->ncols[r], p1->nrows[r],
                              p->window_width, p->window_height, p->step_factor, p->max_iterations,
                              p->min_determinant, p->min_displacement, p->max_residue,
                              p->lighting_insensitive);
      if (v == KLTO_SMALL_DET || v == KLTO_OOB) break;
    }
    if (v == KLTO_OOB ||
        xout < p->borderx || xout > ncols - 1 - p->borderx ||
        yout < p->bordery || yout > nrows - 1 - p->bordery) {
      x[f] = -1.0f; y[f] = -1.0f; val[f] = KLTO_OOB; st[f].has = 0;
    } else if (v == KLTO_SMALL_DET || v == KLTO_LARGE_RESIDUE || v == KLTO_MAX_ITERATIONS) {
      x[f] = -1.0f; y[f] = -1.0f; val[f] = v; st[f].has = 0;
    } else {
      x[f] = xout; y[f] = yout; val[f] = KLTO_TRACKED;
      if (ap->check < 0) continue;
      if (!st[f].has) {
        /* :1446-1457 save the template at the finest level after the first successful track */
        const int hw = tw / 2, hh = th / 2, x0 = (int)xloc, y0 = (int)yloc;
        int k = 0;
        for (j = -hh; j <= hh; j++)
          for (i = -hw; i <= hw; i++, k++) {
            const long off = (long)(j + y0) * ncols + (i + x0);
            t_img[k] = p1->img[0][off]; t_gx[k] = p1->gx[0][off]; t_gy[k] = p1->gy[0][off];
          }
        st[f].aff_x = xloc - (int)xloc + (ap->window_width + 2) / 2;
        st[f].aff_y = yloc - (int)yloc + (ap->window_height + 2) / 2;
        st[f].has = 1;
      } else {
        /* :1458-1493 */
        float xo = xout, yo = yout;
        v = klto_track_affine_feature(st[f].aff_x, st[f].aff_y, &xo, &yo, t_img, t_gx, t_gy, tw, th,
                                      p2->img[0], p2->gx[0], p2->gy[0], ncols, nrows,
                                      ap->window_width, ap->window_height, p->step_factor,
                                      ap->max_iterations, p->min_determinant, p->min_displacement,
                                      ap->min_displacement, ap->max_residue, ap->check,
                                      ap->max_displacement_differ,
                                      &st[f].Axx, &st[f].Ayx, &st[f].Axy, &st[f].Ayy,
                                      p->lighting_insensitive);
        val[f] = v;
        if (v != KLTO_TRACKED) {
          x[f] = -1.0f; y[f] = -1.0f; st[f].aff_x = -1.0f; st[f].aff_y = -1.0f; st[f].has = 0;
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------ */
/* selection -- reference selectGoodFeatures.c                              */
/* ------------------------------------------------------------------------ */

/* reference selectGoodFeatures.c:289-292 (_minEigenvalue): the radicand is a
 * float expression, sqrt and the outer subtraction/division are double. */
static float min_eigenvalue(float gxx, float gxy, float gyy)
{
  return (float)((gxx + gyy - sqrt((gxx - gyy) * (gxx - gyy) + 4 * gxy * gxy)) / 2.0f);
}

/* reference selectGoodFeatures.c:373-424 */
int klto_mineig_points(const float *gx, const float *gy, int ncols, int nrows,
                       int ww, int wh, int borderx, int bordery, int skip,
                       int *points)
{
  const int hw = ww / 2, hh = wh / 2;
  int x, y, xx, yy, n = 0;
  if (borderx < hw) borderx = hw;
  if (bordery < hh) bordery = hh;
  for (y = bordery; y < nrows - bordery; y += skip + 1)
    for (x = borderx; x < ncols - borderx; x += skip + 1) {
      float gxx = 0, gxy = 0, gyy = 0;
      for (yy = y - hh; yy <= y + hh; yy++)
        for (xx = x - hw; xx <= x + hw; xx++) {
          const float a = gx[(long)ncols * yy + xx];
          const float b = gy[(long)ncols * yy + xx];
          gxx += a * a;
          gxy += a * b;
          gyy += b * b;
        }
      /* the reference clamps to `limit`, which overflows to UINT_MAX
       * (selectGoodFeatures.c:381,391-392) -- the clamp never fires. */
      points[3 * n + 0] = x;
      points[3 * n + 1] = y;
      points[3 * n + 2] = (int)min_eigenvalue(gxx, gxy, gyy);
      n++;
    }
  return n;
}

/* -- sorting --------------------------------------------------------------- */

static void swap3(int *p, unsigned a, unsigned b)
{
  int t0 = p[3 * a], t1 = p[3 * a + 1], t2 = p[3 * a + 2];
  p[3 * a] = p[3 * b]; p[3 * a + 1] = p[3 * b + 1]; p[3 * a + 2] = p[3 * b + 2];
  p[3 * b] = t0; p[3 * b + 1] = t1; p[3 * b + 2] = t2;
}

/* reference selectGoodFeatures.c:62-96 (_quicksort).  The permutation it
 * produces for equal keys is part of the reference's observable behaviour
 * (SURVEY.md 7-H1), so the partition scheme is restated step for step:
 * middle element to the front as pivot, two inward scans, pivot to its final
 * place, recurse into the smaller side and loop on the larger. */
static void quick_desc(int *p, int n)
{
  while (n > 1) {
    unsigned i = 0, j = (unsigned)n, left, right;
    swap3(p, 0, (unsigned)(n / 2));
    for (;;) {
      do { --j; } while (p[3 * j + 2] < p[2]);
      do { ++i; } while (i < j && p[3 * i + 2] > p[2]);
      if (i >= j) break;
      swap3(p, i, j);
    }
    swap3(p, j, 0);
    left = j;
    right = (unsigned)n - (j + 1);
    j++;
    if (left < right) {
      quick_desc(p, (int)left);
      p += 3 * j;
      n = (int)right;
    } else {
      quick_desc(p + 3 * j, (int)right);
      n = (int)left;
    }
  }
}

/* stable descending merge sort == glibc qsort with the reference's
 * _comparePoints (selectGoodFeatures.c:250-260) under -DKLT_USE_QSORT */
static void merge_desc(int *p, int *tmp, int n)
{
  int half, i, j, k;
  if (n < 2) return;
  half = n / 2;
  merge_desc(p, tmp, half);
  merge_desc(p + 3 * half, tmp, n - half);
  i = 0; j = half; k = 0;
  while (i < half && j < n) {
    const int *src = (p[3 * j + 2] > p[3 * i + 2]) ? &p[3 * j++] : &p[3 * i++];
    tmp[3 * k] = src[0]; tmp[3 * k + 1] = src[1]; tmp[3 * k + 2] = src[2]; k++;
  }
  while (i < half) { memcpy(tmp + 3 * k, p + 3 * i, 3 * sizeof(int)); i++; k++; }
  while (j < n)    { memcpy(tmp + 3 * k, p + 3 * j, 3 * sizeof(int)); j++; k++; }
  memcpy(p, tmp, sizeof(int) * 3 * (size_t)n);
}

void klto_sort_points(int *points, int npoints, int sort_kind)
{
  if (sort_kind == KLTO_SORT_QUICK) {
    quick_desc(points, npoints);
  } else {
    int *tmp = (int *)malloc(sizeof(int) * 3 * (size_t)(npoints > 0 ? npoints : 1));
    merge_desc(points, tmp, npoints);
    free(tmp);
  }
}

/* reference selectGoodFeatures.c:102-115 (_fillFeaturemap) */
static void stamp(unsigned char *map, int x, int y, int d, int ncols, int nrows)
{
  int ix, iy;
  for (iy = y - d; iy <= y + d; iy++)
    for (ix = x - d; ix <= x + d; ix++)
      if (ix >= 0 && ix < ncols && iy >= 0 && iy < nrows) map[(long)iy * ncols + ix] = 1;
}

/* reference selectGoodFeatures.c:135-239 (_enforceMinimumDistance) */
void klto_enforce_min_distance(const int *points, int npoints, int ncols, int nrows,
                               int mindist, int min_eigenvalue, int overwrite_all,
                               int n, float *x, float *y, int *val)
{
  unsigned char *map = (unsigned char *)calloc((size_t)ncols * nrows, 1);
  int slot = 0, c = 0, i;
  if (min_eigenvalue < 1) min_eigenvalue = 1;
  mindist--;
  if (!overwrite_all)
    for (i = 0; i < n; i++)
      if (val[i] >= 0) stamp(map, (int)x[i], (int)y[i], mindist, ncols, nrows);

  for (;;) {
    int px, py, pv;
    if (c >= npoints) {
      /* ran out of candidates: every slot still open becomes NOT_FOUND */
      for (; slot < n; slot++)
        if (overwrite_all || val[slot] < 0) { x[slot] = -1; y[slot] = -1; val[slot] = KLTO_NOT_FOUND; }
      break;
    }
    px = points[3 * c]; py = points[3 * c + 1]; pv = points[3 * c + 2];
    c++;
    while (!overwrite_all && slot < n && val[slot] >= 0) slot++;
    if (slot >= n) break;
    if (!map[(long)py * ncols + px] && pv >= min_eigenvalue) {
      x[slot] = (float)px; y[slot] = (float)py; val[slot] = pv;
      slot++;
      stamp(map, px, py, mindist, ncols, nrows);
    }
  }
  free(map);
}

/* reference selectGoodFeatures.c:297-453 (_KLTSelectGoodFeatures) */
void klto_select(const unsigned char *img, int ncols, int nrows,
                 const klto_pyramids *last, const klto_params *pin,
                 int sort_kind, int replace, int n, float *x, float *y, int *val)
{
  klto_params p = *pin;
  size_t npx = (size_t)ncols * nrows;
  float *fimg = NULL, *gx = NULL, *gy = NULL;
  const float *cgx, *cgy;
  int *points = (int *)malloc(sizeof(int) * 3 * npx);
  int npoints;

  fix_window(&p);
  if (replace && last != NULL) {
    cgx = last->gx[0]; cgy = last->gy[0];
  } else {
    float *tmp = (float *)malloc(sizeof(float) * npx);
    fimg = (float *)malloc(sizeof(float) * npx);
    gx = (float *)malloc(sizeof(float) * npx);
    gy = (float *)malloc(sizeof(float) * npx);
    if (p.smoothBeforeSelecting) {
      klto_to_float(img, ncols, nrows, tmp);
      klto_smooth(tmp, ncols, nrows, klto_smooth_sigma(&p), fimg);
    } else {
      klto_to_float(img, ncols, nrows, fimg);
    }
    klto_gradients(fimg, ncols, nrows, p.grad_sigma, gx, gy);
    free(tmp);
    cgx = gx; cgy = gy;
  }
  npoints = klto_mineig_points(cgx, cgy, ncols, nrows, p.window_width, p.window_height,
                               p.borderx, p.bordery, p.nSkippedPixels, points);
  klto_sort_points(points, npoints, sort_kind);
  if (p.mindist < 0) p.mindist = 0;
  klto_enforce_min_distance(points, npoints, ncols, nrows, p.mindist, p.min_eigenvalue,
                            !replace, n, x, y, val);
  free(points); free(fimg); free(gx); free(gy);
}
