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

/* Batched driver API: track fl through a whole sequence of HOST frames in one call.
 * Equivalent, result for result, to the reference's driver loop (src/V3/example3.c:54-76)
 *
 *     for (k = 1; k < nframes; k++) {
 *       KLTTrackFeatures(tc, frames[k-1], frames[k], ncols, nrows, fl);
 *       if (replace) KLTReplaceLostFeatures(tc, frames[k], ncols, nrows, fl);
 *       if (ft) KLTStoreFeatureList(fl, ft, first_frame + k);
 *     }
 *
 * with tc->sequentialMode = TRUE, but nothing waits for a frame's result before the next frame is
 * queued: frame k+1 crosses PCIe (two device staging buffers) while frame k's pyramids are built
 * and its features tracked, the features stay in HBM, every frame's x | y | val snapshot is copied
 * into pinned memory behind the tracker, and the host synchronises once at the end.  frames[0] is
 * only read when tc holds no previous pyramid (as img1 in KLTTrackFeatures) and may then not be
 * NULL.  ft may be NULL; column first_frame + k receives frame k (column first_frame is not
 * written: the caller stores the selection itself, as example3 does).  On return fl holds the
 * state after the last frame and tc is in sequential mode with the last frame's pyramids held.
 * tc->affineConsistencyCheck >= 0 is honoured (reference trackFeatures.c:1438-1497): the per-feature
 * affine state and templates live on the device during the call and are attached to fl at its end,
 * exactly as the loop above leaves them. */
void KLTTrackFeaturesSequence(KLT_TrackingContext tc, KLT_PixelType *const *frames, int nframes,
                              int ncols, int nrows, KLT_FeatureList fl, KLT_FeatureTable ft,
                              int first_frame, int replace);

#ifdef __cplusplus
}
#endif
#endif
