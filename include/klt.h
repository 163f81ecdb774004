/* include/klt.h -- public C API of the B200-native KLT tracker.
 *
 * Drop-in for the reference's klt.h (reference src/V4/klt.h:14-233; V2/V3 add
 * the extern "C" guards kept here).  Struct layouts, type names, status codes
 * and the 29 prototypes are ABI-identical so that the reference's example3
 * drivers (src/V1/example3.c, src/V3/example3.c) compile and link unchanged.
 * Layout check: sizeof(KLT_FeatureRec) == 64, sizeof(KLT_TrackingContextRec)
 * == 136 on LP64 (tests/test_host.py::test_abi_matches_c_compiler).
 *
 * What differs is below the API: KLTSelectGoodFeatures, KLTTrackFeatures and
 * KLTReplaceLostFeatures run on the GPU (include/klt_cuda.h); there is no CPU
 * implementation of that path in this library.
 */
#ifndef _KLT_H_
#define _KLT_H_

#ifdef __cplusplus
extern "C" {
#endif

typedef float KLT_locType;
typedef unsigned char KLT_PixelType;

#define KLT_BOOL int

#ifndef TRUE
#define TRUE  1
#define FALSE 0
#endif

#ifndef NULL
#define NULL  0
#endif

/* feature status: val > 0 freshly selected (its eigenvalue), 0 tracked, < 0 lost */
#define KLT_TRACKED           0
#define KLT_NOT_FOUND        -1
#define KLT_SMALL_DET        -2
#define KLT_MAX_ITERATIONS   -3
#define KLT_OOB              -4
#define KLT_LARGE_RESIDUE    -5

/* dense row-major float image (reference src/V4/klt_util.h:8-12); only used by
 * the aff_img* members below: the per-feature templates of the affine consistency check, mirrored
 * from the device store (csrc/klt_affine.cuh) when tc->affineConsistencyCheck >= 0. */
#ifndef _KLT_UTIL_H_
#define _KLT_UTIL_H_
typedef struct {
  int ncols;
  int nrows;
  float *data;
} _KLT_FloatImageRec, *_KLT_FloatImage;
#endif

typedef struct {
  /* user-settable */
  int mindist;                    /* minimum distance between selected features */
  int window_width, window_height;
  KLT_BOOL sequentialMode;        /* keep the last frame's pyramids between calls */
  KLT_BOOL smoothBeforeSelecting;
  KLT_BOOL writeInternalImages;   /* accepted, ignored (debug dumps are out of scope) */
  KLT_BOOL lighting_insensitive;  /* TRUE: gain / bias normalised windows (reference trackFeatures.c:125-220) */

  int min_eigenvalue;
  float min_determinant;
  float min_displacement;
  int max_iterations;
  float max_residue;
  float grad_sigma;
  float smooth_sigma_fact;
  float pyramid_sigma_fact;
  float step_factor;
  int nSkippedPixels;
  int borderx;
  int bordery;
  int nPyramidLevels;
  int subsampling;

  /* affine consistency check (reference trackFeatures.c:1438-1497): -1 off, 0 translation,
   * 1 similarity, 2 affine -- runs on the device (csrc/klt_affine.cuh) */
  int affine_window_width, affine_window_height;
  int affineConsistencyCheck;
  int affine_max_iterations;
  float affine_max_residue;
  float affine_min_displacement;
  float affine_max_displacement_differ;

  /* library-owned.  Non-NULL exactly when the previous frame's pyramids are
   * held (on the device) for sequentialMode, as in the reference. */
  void *pyramid_last;
  void *pyramid_last_gradx;
  void *pyramid_last_grady;
} KLT_TrackingContextRec, *KLT_TrackingContext;

typedef struct {
  KLT_locType x;
  KLT_locType y;
  int val;
  _KLT_FloatImage aff_img;
  _KLT_FloatImage aff_img_gradx;
  _KLT_FloatImage aff_img_grady;
  KLT_locType aff_x;
  KLT_locType aff_y;
  KLT_locType aff_Axx;
  KLT_locType aff_Ayx;
  KLT_locType aff_Axy;
  KLT_locType aff_Ayy;
} KLT_FeatureRec, *KLT_Feature;

typedef struct {
  int nFeatures;
  KLT_Feature *feature;
} KLT_FeatureListRec, *KLT_FeatureList;

typedef struct {
  int nFrames;
  KLT_Feature *feature;
} KLT_FeatureHistoryRec, *KLT_FeatureHistory;

typedef struct {
  int nFrames;
  int nFeatures;
  KLT_Feature **feature;
} KLT_FeatureTableRec, *KLT_FeatureTable;

/* create / free */
KLT_TrackingContext KLTCreateTrackingContext(void);
KLT_FeatureList KLTCreateFeatureList(int nFeatures);
KLT_FeatureHistory KLTCreateFeatureHistory(int nFrames);
KLT_FeatureTable KLTCreateFeatureTable(int nFrames, int nFeatures);
void KLTFreeTrackingContext(KLT_TrackingContext tc);
void KLTFreeFeatureList(KLT_FeatureList fl);
void KLTFreeFeatureHistory(KLT_FeatureHistory fh);
void KLTFreeFeatureTable(KLT_FeatureTable ft);

/* the hot path (GPU) */
void KLTSelectGoodFeatures(KLT_TrackingContext tc, KLT_PixelType *img,
                           int ncols, int nrows, KLT_FeatureList fl);
void KLTTrackFeatures(KLT_TrackingContext tc, KLT_PixelType *img1,
                      KLT_PixelType *img2, int ncols, int nrows,
                      KLT_FeatureList fl);
void KLTReplaceLostFeatures(KLT_TrackingContext tc, KLT_PixelType *img,
                            int ncols, int nrows, KLT_FeatureList fl);

/* utilities */
int KLTCountRemainingFeatures(KLT_FeatureList fl);
void KLTPrintTrackingContext(KLT_TrackingContext tc);
void KLTChangeTCPyramid(KLT_TrackingContext tc, int search_range);
void KLTUpdateTCBorder(KLT_TrackingContext tc);
void KLTStopSequentialMode(KLT_TrackingContext tc);
void KLTSetVerbosity(int verbosity);
float _KLTComputeSmoothSigma(KLT_TrackingContext tc);

/* feature list <-> table */
void KLTStoreFeatureList(KLT_FeatureList fl, KLT_FeatureTable ft, int frame);
void KLTExtractFeatureList(KLT_FeatureList fl, KLT_FeatureTable ft, int frame);
void KLTStoreFeatureHistory(KLT_FeatureHistory fh, KLT_FeatureTable ft, int feat);
void KLTExtractFeatureHistory(KLT_FeatureHistory fh, KLT_FeatureTable ft, int feat);

/* files */
void KLTWriteFeatureListToPPM(KLT_FeatureList fl, KLT_PixelType *greyimg,
                              int ncols, int nrows, char *filename);
void KLTWriteFeatureList(KLT_FeatureList fl, char *filename, char *fmt);
void KLTWriteFeatureHistory(KLT_FeatureHistory fh, char *filename, char *fmt);
void KLTWriteFeatureTable(KLT_FeatureTable ft, char *filename, char *fmt);
KLT_FeatureList KLTReadFeatureList(KLT_FeatureList fl, char *filename);
KLT_FeatureHistory KLTReadFeatureHistory(KLT_FeatureHistory fh, char *filename);
KLT_FeatureTable KLTReadFeatureTable(KLT_FeatureTable ft, char *filename);

/* errors (reference src/V4/error.h): KLTError prints "KLT Error: ..." and
 * exit(1)s; KLTWarning prints and returns. */
void KLTError(char *fmt, ...);
void KLTWarning(char *fmt, ...);

#ifdef __cplusplus
}
#endif

#endif
