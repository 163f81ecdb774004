/* oracle/klt_oracle.h
 *
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference's CPU algorithm for the KLT hot path
 * (select -> pyramid/gradients -> track).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this.  The
 * product library (libklt_b200.so) never links, loads or calls it.
 *
 * Parity status: PINNED.  liboracle.so is checked (tests/test_oracle.py)
 *   - byte-for-byte against the reference's checked-in golden run
 *     src/V1/feat/features2.ft (copied to tests/golden/features2.ft), and
 *   - bit-for-bit, stage by stage, against oracle/_ref/libklt_ref.so, which is
 *     the unmodified reference CPU sources (the .c files of src/V3) compiled in place.
 *
 * All arithmetic is IEEE float32, round-to-nearest, no FMA contraction, in the
 * summation order of the reference (file:line cited at each function).
 */
#ifndef KLT_ORACLE_H
#define KLT_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define KLTO_MAX_TAPS 71

#define KLTO_TRACKED         0
#define KLTO_NOT_FOUND      -1
#define KLTO_SMALL_DET      -2
#define KLTO_MAX_ITERATIONS -3
#define KLTO_OOB            -4
#define KLTO_LARGE_RESIDUE  -5

#define KLTO_SORT_STABLE    0   /* == reference built with -DKLT_USE_QSORT on glibc */
#define KLTO_SORT_QUICK     1   /* == reference default build (_quicksort)           */

typedef struct {
  int   mindist;
  int   window_width, window_height;
  int   smoothBeforeSelecting;
  int   min_eigenvalue;
  float min_determinant;
  float min_displacement;
  int   max_iterations;
  float max_residue;
  float grad_sigma;
  float smooth_sigma_fact;
  float pyramid_sigma_fact;
  float step_factor;
  int   nSkippedPixels;
  int   borderx, bordery;
  int   nPyramidLevels;
  int   subsampling;
  int   lighting_insensitive;   /* tc->lighting_insensitive (klt.h:50), default FALSE */
} klto_params;

/* parameter derivation (klt.c:20-44, :288-343, :362-431) */
void klto_default_params(klto_params *p);
void klto_change_pyramid(klto_params *p, int search_range);
void klto_update_border(klto_params *p);
float klto_smooth_sigma(const klto_params *p);

/* taps (convolve.c:60-130); returns 0 on success, -1 if sigma needs > 71 taps */
int klto_taps(float sigma, float *gauss, int *gauss_width,
              float *deriv, int *deriv_width);
/* forget the cached sigma (the reference keeps one in a file-static) */
void klto_reset_tap_cache(void);

/* stages (dense row-major float images, no pitch) */
void klto_to_float(const unsigned char *img, int ncols, int nrows, float *out);
void klto_convolve_separate(const float *in, int ncols, int nrows,
                            const float *kh, int wh, const float *kv, int wv,
                            float *out);
void klto_smooth(const float *in, int ncols, int nrows, float sigma, float *out);
void klto_gradients(const float *in, int ncols, int nrows, float sigma,
                    float *gx, float *gy);
/* one pyramid step: out has (ncols/ss) x (nrows/ss) pixels */
void klto_pyr_down(const float *in, int ncols, int nrows, int ss, float sigma,
                   float *out);

/* eigenvalue map: fills triples (x,y,val) in raster order, returns count */
int klto_mineig_points(const float *gx, const float *gy, int ncols, int nrows,
                       int ww, int wh, int borderx, int bordery, int skip,
                       int *points);

/* pyramids of one frame */
typedef struct klto_pyramids klto_pyramids;
klto_pyramids *klto_build_pyramids(const unsigned char *img, int ncols, int nrows,
                                   const klto_params *p);
void klto_free_pyramids(klto_pyramids *q);
int  klto_pyr_levels(const klto_pyramids *q);
void klto_pyr_dims(const klto_pyramids *q, int level, int *ncols, int *nrows);
/* which: 0 = image, 1 = gradx, 2 = grady */
const float *klto_pyr_data(const klto_pyramids *q, int which, int level);

/* KLTTrackFeatures body for one frame pair (trackFeatures.c:1343-1437) */
void klto_track(const klto_pyramids *p1, const klto_pyramids *p2,
                const klto_params *p, int n, float *x, float *y, int *val);

/* affine consistency check (trackFeatures.c:506-1224, :1438-1497; klt.c:33-39 defaults) */
typedef struct {
  int   check;                  /* tc->affineConsistencyCheck: -1 off, 0 translation, 1 similarity, 2 affine */
  int   window_width, window_height;     /* affine_window_* (15) */
  int   max_iterations;         /* affine_max_iterations (10) */
  float max_residue;            /* affine_max_residue (10) */
  float min_displacement;       /* affine_min_displacement (0.02) */
  float max_displacement_differ;/* affine_max_displacement_differ (1.5) */
} klto_affine_params;
typedef struct {                /* the aff_* members of KLT_FeatureRec (klt.h:97-105) */
  int   has;                    /* aff_img != NULL */
  float aff_x, aff_y, Axx, Ayx, Axy, Ayy;
} klto_affine_state;
void klto_track_affine(const klto_pyramids *p1, const klto_pyramids *p2, const klto_params *p,
                       const klto_affine_params *ap, int n, float *x, float *y, int *val,
                       klto_affine_state *st, float *tmpl);
int klto_track_affine_feature(float x1, float y1, float *x2, float *y2,
                              const float *img1, const float *gx1, const float *gy1, int nc1, int nr1,
                              const float *img2, const float *gx2, const float *gy2, int nc2, int nr2,
                              int width, int height, float step_factor, int max_iterations,
                              float small, float th, float th_aff, float max_residue,
                              int affine_map, float mdd,
                              float *Axx, float *Ayx, float *Axy, float *Ayy, int lighting);

/* single level solver, exposed for edge-case tests (trackFeatures.c:381-486) */
int klto_track_level(float x1, float y1, float *x2, float *y2,
                     const float *img1, const float *gx1, const float *gy1,
                     const float *img2, const float *gx2, const float *gy2,
                     int ncols, int nrows, int ww, int wh, float step_factor,
                     int max_iterations, float small, float th, float max_residue);

/* selection (selectGoodFeatures.c:297-453).
 * replace == 0: KLTSelectGoodFeatures (overwrite all)
 * replace == 1: KLTReplaceLostFeatures; if last != NULL its level-0 image and
 *               gradients are used and img is ignored (sequentialMode path). */
void klto_select(const unsigned char *img, int ncols, int nrows,
                 const klto_pyramids *last, const klto_params *p,
                 int sort_kind, int replace, int n, float *x, float *y, int *val);

/* greedy minimum-distance pass on an already sorted triple list */
void klto_enforce_min_distance(const int *points, int npoints, int ncols, int nrows,
                               int mindist, int min_eigenvalue, int overwrite_all,
                               int n, float *x, float *y, int *val);
void klto_sort_points(int *points, int npoints, int sort_kind);

#ifdef __cplusplus
}
#endif
#endif
