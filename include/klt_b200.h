/* include/klt_b200.h -- extensions of the KLT C API that only make sense for
 * the GPU library.  Nothing here is needed by the reference's drivers; they
 * exist for multi-GPU placement, parity testing and pipelined benches. */
#ifndef KLT_B200_H
#define KLT_B200_H

#include <stddef.h>
#include "klt.h"
#include "klt_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* CUDA device the context's work runs on.  Default: environment variable
 * KLT_B200_DEVICE, else the calling thread's current device.  Must be called
 * before the first Select/Track call on tc. */
void KLTB200SetDevice(KLT_TrackingContext tc, int device);

/* 1: tracking pyramids and the tracker use reference-order, separately rounded
 * multiply/add (bit-identical to the CPU reference); 0 (default, or env
 * KLT_B200_EXACT unset/0): fused multiply-add, within 1e-4 relative.
 * KLTSelectGoodFeatures always builds its images in exact mode, because its
 * ranking keys are truncated integers. */
void KLTB200SetExact(KLT_TrackingContext tc, int exact);
int  KLTB200GetExact(KLT_TrackingContext tc);

/* the device context behind tc (created on demand) */
klt_dev *KLTB200Device(KLT_TrackingContext tc);
/* slot (0/1) holding tc's previous-frame pyramids, -1 if none */
int KLTB200LastSlot(KLT_TrackingContext tc);

/* KLTTrackFeatures with the new frame already resident in HBM (row pitch in
 * bytes).  img1 is only used when no previous pyramid is held, as in the
 * reference (src/V1/trackFeatures.c:1285-1308); it may be NULL otherwise. */
void KLTTrackFeaturesDevice(KLT_TrackingContext tc, const KLT_PixelType *d_img1,
                            const KLT_PixelType *d_img2, size_t pitch,
                            int ncols, int nrows, KLT_FeatureList fl);

/* Resident-feature tracking for pipelined drivers: the feature arrays stay on
 * the device between frames, no host synchronisation per frame.
 *   KLTB200ResidentBegin  uploads fl and (if needed) builds the first pyramid
 *   KLTB200ResidentStep   builds pyramids of one more frame and tracks into it
 *   KLTB200ResidentEnd    downloads the features into fl (synchronises)
 * frames are host pointers when on_device == 0 (pinned memory recommended). */
void KLTB200ResidentBegin(KLT_TrackingContext tc, const KLT_PixelType *img1, int on_device,
                          size_t pitch, int ncols, int nrows, KLT_FeatureList fl);
void KLTB200ResidentStep(KLT_TrackingContext tc, const KLT_PixelType *img2, int on_device,
                         size_t pitch, int ncols, int nrows);
void KLTB200ResidentEnd(KLT_TrackingContext tc, KLT_FeatureList fl);

#ifdef __cplusplus
}
#endif
#endif
