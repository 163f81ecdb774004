/* examples/example3.c -- sequence driver on the public KLT API (include/klt.h).
 *
 * Same flow as the reference's drivers (select on the first frame, track frame to
 * frame in sequentialMode, optionally replace lost features, store every frame in a
 * feature table, write the table), with the paths taken from the command line:
 *
 *   example3 <dir> <first> <nFrames> <nFeatures> [out_prefix] [replace]
 *
 * reads <dir>/img<first>.pgm ... and times KLTTrackFeatures only, like the
 * reference's src/V3/example3.c:61-65.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "klt.h"
#include "pnmio.h"

static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char **argv)
{
  char name[1024];
  unsigned char *img1, *img2;
  KLT_TrackingContext tc;
  KLT_FeatureList fl;
  KLT_FeatureTable ft;
  int first, nFrames, nFeatures, ncols, nrows, i, replace;
  const char *dir, *prefix;
  double t_track = 0.0;
  long tracked = 0;

  if (argc < 5) {
    fprintf(stderr, "usage: %s <dir> <first> <nFrames> <nFeatures> [out_prefix] [replace]\n", argv[0]);
    return 2;
  }
  dir = argv[1];
  first = atoi(argv[2]);
  nFrames = atoi(argv[3]);
  nFeatures = atoi(argv[4]);
  prefix = argc > 5 ? argv[5] : "features";
  replace = argc > 6 ? atoi(argv[6]) : 0;

  tc = KLTCreateTrackingContext();
  fl = KLTCreateFeatureList(nFeatures);
  ft = KLTCreateFeatureTable(nFrames, nFeatures);
  tc->sequentialMode = TRUE;
  tc->writeInternalImages = FALSE;
  tc->affineConsistencyCheck = -1;

  snprintf(name, sizeof name, "%s/img%d.pgm", dir, first);
  img1 = pgmReadFile(name, NULL, &ncols, &nrows);
  img2 = (unsigned char *)malloc((size_t)ncols * nrows);

  KLTSelectGoodFeatures(tc, img1, ncols, nrows, fl);
  KLTStoreFeatureList(fl, ft, 0);

  for (i = 1; i < nFrames; i++) {
    double t0;
    snprintf(name, sizeof name, "%s/img%d.pgm", dir, first + i);
    pgmReadFile(name, img2, &ncols, &nrows);
    tracked += KLTCountRemainingFeatures(fl);
    t0 = now_s();
    KLTTrackFeatures(tc, img1, img2, ncols, nrows, fl);
    t_track += now_s() - t0;
    if (replace) KLTReplaceLostFeatures(tc, img2, ncols, nrows, fl);
    KLTStoreFeatureList(fl, ft, i);
    memcpy(img1, img2, (size_t)ncols * nrows);
  }

  snprintf(name, sizeof name, "%s.txt", prefix);
  KLTWriteFeatureTable(ft, name, "%5.1f");
  snprintf(name, sizeof name, "%s.ft", prefix);
  KLTWriteFeatureTable(ft, name, NULL);

  printf("frames %d  features %d  tracking time %.6f s  (%.1f frames/s, %.0f features/s)\n",
         nFrames, nFeatures, t_track, (nFrames - 1) / t_track, tracked / t_track);

  KLTFreeFeatureTable(ft);
  KLTFreeFeatureList(fl);
  KLTFreeTrackingContext(tc);
  free(img1);
  free(img2);
  return 0;
}
