/* klt_select.c -- KLTSelectGoodFeatures / KLTReplaceLostFeatures on the GPU.
 *
 * Host orchestration of reference src/V1/selectGoodFeatures.c:297-541:
 * window repair (:313-333), choice of the image source (:342-364: in
 * REPLACING_SOME + sequentialMode the level-0 gradients of pyramid_last are
 * reused and `img` is ignored), the eigenvalue map + ranking + minimum-distance
 * pass (device, csrc/klt_dev.cu) and the verbose messages (:478-494, :523-540).
 *
 * Ranking ties are broken in raster order (a stable sort of the raster-ordered
 * candidates); the default reference build uses an unstable hand-written
 * quicksort whose tie order a parallel sort cannot reproduce, its
 * -DKLT_USE_QSORT build happens to sort stably with glibc.
 */
#include <stdio.h>
#include <stdlib.h>

#include "klt_internal.h"

#define DEVCALL(s, call)                                                     \
  do {                                                                       \
    if ((call) != 0) KLTError("(KLT/B200) %s", klt_dev_error((s)->dev));      \
  } while (0)

/* parameters of the device selection, with the reference's repair of a negative mindist
 * (selectGoodFeatures.c:430-434) */
void klt_fill_select_params(KLT_TrackingContext tc, int replacing, klt_dev_select_params *sp)
{
  if (tc->mindist < 0) {
    KLTWarning("(_KLTSelectGoodFeatures) Tracking context field tc->mindist "
               "is negative (%d); setting to zero", tc->mindist);
    tc->mindist = 0;
  }
  sp->window_width = tc->window_width;
  sp->window_height = tc->window_height;
  sp->borderx = tc->borderx;
  sp->bordery = tc->bordery;
  sp->nSkippedPixels = tc->nSkippedPixels;
  sp->mindist = tc->mindist;
  sp->min_eigenvalue = tc->min_eigenvalue;
  sp->overwrite_all = replacing ? 0 : 1;
}

static void select_common(KLT_TrackingContext tc, const KLT_PixelType *img, int ncols, int nrows,
                          KLT_FeatureList fl, int replacing)
{
  klt_tc_state *s = klt_state_get(tc);
  klt_dev *dev = klt_state_device(s);
  klt_dev_select_params sp;
  const int n = fl->nFeatures;
  int slot, i;
  float *x, *y;
  int *v;

  klt_fix_window(tc, "KLTSelectGoodFeatures", 1);

  if (replacing && tc->sequentialMode && tc->pyramid_last != NULL && s->last_slot >= 0 &&
      klt_dev_slot_valid(dev, s->last_slot)) {
    int gw = 0, gh = 0;
    klt_dev_geometry(dev, &gw, &gh, NULL, NULL);
    if (gw != ncols || gh != nrows)
      KLTError("(KLTReplaceLostFeatures) image is %d by %d but the stored pyramid is %d by %d",
               ncols, nrows, gw, gh);
    /* reuse, img ignored (:342-348).  The ranking keys are truncated integers, so level 0 must
     * be in exact arithmetic: a slot tracked in fma mode gets its level 0 rebuilt (exact) from the
     * frame still on the device */
    DEVCALL(s, klt_dev_exact_level0(dev, s->last_slot, &slot));
  } else {
    /* level 0 + its gradients only; always exact arithmetic so the integer
     * eigenvalues equal the CPU reference's.  Built in the slot that does not
     * hold the previous frame. */
    klt_dev_build_desc q;
    slot = (tc->pyramid_last != NULL && s->last_slot >= 0) ? (s->last_slot + 1) % KLT_DEV_SLOTS : 0;
    klt_fill_build_desc(tc, ncols, nrows, 1, tc->smoothBeforeSelecting ? 1 : 0, 1, &q);
    DEVCALL(s, klt_dev_build(dev, slot, img, 0, (size_t)ncols, &q));
  }

  klt_fill_select_params(tc, replacing, &sp);

  x = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  y = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  v = (int *)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
  if (!x || !y || !v) KLTError("(KLTSelectGoodFeatures) Out of memory");
  klt_list_to_arrays(fl, x, y, v);
  DEVCALL(s, klt_dev_select(dev, slot, &sp, n, x, y, v));
  for (i = 0; i < n; i++) {
    KLT_Feature f = fl->feature[i];
    if (replacing && f->val >= 0) continue;       /* kept features are not rewritten */
    f->x = x[i];
    f->y = y[i];
    f->val = v[i];
    f->aff_img = NULL;
    f->aff_img_gradx = NULL;
    f->aff_img_grady = NULL;
    f->aff_x = -1.0f;
    f->aff_y = -1.0f;
    f->aff_Axx = 1.0f;
    f->aff_Ayx = 0.0f;
    f->aff_Axy = 0.0f;
    f->aff_Ayy = 1.0f;
  }
  free(x); free(y); free(v);
}

void KLTSelectGoodFeatures(KLT_TrackingContext tc, KLT_PixelType *img, int ncols, int nrows,
                           KLT_FeatureList fl)
{
  if (KLT_verbose >= 1) {
    fprintf(stderr, "(KLT) Selecting the %d best features from a %d by %d image...  ",
            fl->nFeatures, ncols, nrows);
    fflush(stderr);
  }
  select_common(tc, img, ncols, nrows, fl, 0);
  if (KLT_verbose >= 1) {
    fprintf(stderr, "\n\t%d features found.\n", KLTCountRemainingFeatures(fl));
    fflush(stderr);
  }
}

void KLTReplaceLostFeatures(KLT_TrackingContext tc, KLT_PixelType *img, int ncols, int nrows,
                            KLT_FeatureList fl)
{
  const int lost = fl->nFeatures - KLTCountRemainingFeatures(fl);
  if (KLT_verbose >= 1) {
    fprintf(stderr, "(KLT) Attempting to replace %d features in a %d by %d image...  ",
            lost, ncols, nrows);
    fflush(stderr);
  }
  if (lost > 0) select_common(tc, img, ncols, nrows, fl, 1);
  if (KLT_verbose >= 1) {
    fprintf(stderr, "\n\t%d features replaced.\n",
            lost - fl->nFeatures + KLTCountRemainingFeatures(fl));
    fflush(stderr);
  }
}
