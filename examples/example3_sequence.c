/* examples/example3_sequence.c -- the driver of example3.c on the batched call.
 *
 *   example3_sequence <dir> <first> <nFrames> <nFeatures> [out_prefix] [replace]
 *
 * Reads all frames first (as a driver with a read-ahead thread or a capture ring would have
 * them), selects on the first one and hands the whole sequence to KLTTrackFeaturesSequence
 * (include/klt_b200.h), which replaces the per-frame loop of the reference's driver
 * (src/V3/example3.c:54-76).  Writes the same two feature-table files as example3, byte for
 * byte (tests/test_gpu_dropin.py), and times the one call.
 */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "klt.h"
#include "klt_b200.h"
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
  unsigned char **img;
  KLT_TrackingContext tc;
  KLT_FeatureList fl;
  KLT_FeatureTable ft;
  int first, nFrames, nFeatures, ncols = 0, nrows = 0, i, replace;
  const char *dir, *prefix;
  double t0, t_track;

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

  img = (unsigned char **)malloc(sizeof(*img) * (size_t)nFrames);
  for (i = 0; i < nFrames; i++) {
    snprintf(name, sizeof name, "%s/img%d.pgm", dir, first + i);
    img[i] = pgmReadFile(name, NULL, &ncols, &nrows);
  }

  KLTSelectGoodFeatures(tc, img[0], ncols, nrows, fl);
  KLTStoreFeatureList(fl, ft, 0);

  t0 = now_s();
  KLTTrackFeaturesSequence(tc, img, nFrames, ncols, nrows, fl, ft, 0, replace);
  t_track = now_s() - t0;

  snprintf(name, sizeof name, "%s.txt", prefix);
  KLTWriteFeatureTable(ft, name, "%5.1f");
  snprintf(name, sizeof name, "%s.ft", prefix);
  KLTWriteFeatureTable(ft, name, NULL);

  printf("frames %d  features %d  sequence call %.6f s  (%.1f frames/s)\n",
         nFrames, nFeatures, t_track, (nFrames - 1) / t_track);

  KLTFreeFeatureTable(ft);
  KLTFreeFeatureList(fl);
  KLTFreeTrackingContext(tc);
  for (i = 0; i < nFrames; i++) free(img[i]);
  free(img);
  return 0;
}
