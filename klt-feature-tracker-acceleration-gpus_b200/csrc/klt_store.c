/* klt_store.c -- copy (x, y, val) between feature lists, histories and tables.
 * Behaviour of reference src/V1/storeFeatures.c:15-116, including its range and
 * size checks (fatal through KLTError). */
#include "klt.h"

static void copy3(KLT_Feature dst, const KLT_Feature src)
{
  dst->x = src->x;
  dst->y = src->y;
  dst->val = src->val;
}

static void check_frame(const char *who, KLT_FeatureList fl, KLT_FeatureTable ft, int frame)
{
  if (frame < 0 || frame >= ft->nFrames)
    KLTError("(%s) Frame number %d is not between 0 and %d", who, frame, ft->nFrames - 1);
  if (fl->nFeatures != ft->nFeatures)
    KLTError("(%s) FeatureList and FeatureTable must have the same number of features", who);
}

static void check_feat(const char *who, KLT_FeatureHistory fh, KLT_FeatureTable ft, int feat)
{
  if (feat < 0 || feat >= ft->nFeatures)
    KLTError("(%s) Feature number %d is not between 0 and %d", who, feat, ft->nFeatures - 1);
  if (fh->nFrames != ft->nFrames)
    KLTError("(%s) FeatureHistory and FeatureTable must have the same number of frames", who);
}

void KLTStoreFeatureList(KLT_FeatureList fl, KLT_FeatureTable ft, int frame)
{
  int i;
  check_frame("KLTStoreFeatures", fl, ft, frame);
  for (i = 0; i < fl->nFeatures; i++) copy3(ft->feature[i][frame], fl->feature[i]);
}

void KLTExtractFeatureList(KLT_FeatureList fl, KLT_FeatureTable ft, int frame)
{
  int i;
  check_frame("KLTExtractFeatures", fl, ft, frame);
  for (i = 0; i < fl->nFeatures; i++) copy3(fl->feature[i], ft->feature[i][frame]);
}

void KLTStoreFeatureHistory(KLT_FeatureHistory fh, KLT_FeatureTable ft, int feat)
{
  int i;
  check_feat("KLTStoreFeatureHistory", fh, ft, feat);
  for (i = 0; i < fh->nFrames; i++) copy3(ft->feature[feat][i], fh->feature[i]);
}

void KLTExtractFeatureHistory(KLT_FeatureHistory fh, KLT_FeatureTable ft, int feat)
{
  int i;
  check_feat("KLTExtractFeatureHistory", fh, ft, feat);
  for (i = 0; i < fh->nFrames; i++) copy3(fh->feature[i], ft->feature[feat][i]);
}
