// klt_stream.cuh -- warp-synchronous streaming formulation of the level-0 pipeline (included by
// klt_dev.cu).
//
// l0_stream_kernel: u8 frame -> L0, gx0, gy0, the same results as l0_fused_kernel (reference
// _KLTToFloatImage + _KLTComputeSmoothedImage + _KLTComputeGradients, src/V1/convolve.c:37-314 as
// sequenced by trackFeatures.c:1311-1321), without shared memory and without a single CTA barrier.
//
// ncu on the tile kernel (profiles/r1_v6_ncu_full_summary.csv): issue 53 %, FMA pipe 41 %, shared
// memory 47 % -- no throughput limit is reached; a quarter of the warp samples wait at the four
// per-tile barriers and the stage passes do not divide evenly over the warps.  Here a warp owns a
// strip of 128 columns (4 per lane) and marches down a segment of rows:
//
//   row y   : one coalesced 128 B load of u8 pixels, neighbours' pixels by shuffle of the raw word,
//             horizontal Gaussian hs(y, .) into slot y mod 7 of a register ring
//   row y-2 : L0 = vertical Gaussian over ring slots y-4 .. y            -> stored (float4 per lane)
//             its neighbours' values by 6 shuffles, horizontal DoG / Gaussian into two more rings
//   row y-5 : gx = vertical Gaussian of the DoG ring, gy = vertical DoG of the Gaussian ring -> stored
//
// The row loop is unrolled by 7 so that every ring index is a compile-time constant (the rings are
// registers).  Taps are applied in increasing order, i.e. in the reference's summation order: the
// results are bit-identical to the tile kernels in both arithmetic modes.  (A packed-FFMA2 variant
// of the vertical and 7-tap passes is kept behind KLT_STREAM_FFMA2: measured slower, see sfma2.)
//
// Cost of the formulation: a strip yields 112 of its 128 columns (5-pixel halo, float4 alignment)
// and a segment of HS rows needs 10 warm-up rows.
#pragma once

struct StreamGeo {
  static constexpr int C = 4;                    // columns per lane
  static constexpr int SW = 32 * C;              // strip width
  static constexpr int HALO = 8;                 // unowned columns on either side (>= 5, multiple of 4)
  static constexpr int OWN = SW - 2 * HALO;      // 112 output columns per strip
  static constexpr int RS = 2, RG = 3;
  static constexpr int LAT = RS + RG;            // output row y needs input rows y-5 .. y+5
};

template <bool EXACT>
__device__ __forceinline__ float smac(float acc, float a, float k) { return mac<EXACT>(acc, a, k); }

// (ax, ay) += (vx, vy) * k.  Packed FFMA2 with a uniform-register multiplier pair issues in half the
// slots but runs at 96 FMA/clk/SM; two scalar FFMA with a uniform multiplier run at 116
// (tools/ffma2_probe.cu) -- this kernel is limited by the FMA pipe, not by issue slots.
#ifndef KLT_STREAM_FFMA2
#define KLT_STREAM_FFMA2 0
#endif
__device__ __forceinline__ void sfma2(float& ax, float& ay, float vx, float vy, const TapsF& t, int m) {
#if KLT_STREAM_FFMA2
  ffma2(ax, ay, vx, vy, t.kk[m]);
#else
  ax = fmaf(vx, t.k[m], ax); ay = fmaf(vy, t.k[m], ay);
#endif
}

// Per-warp state that advances by one row per step: the four row pointers (so that no row needs
// an integer multiply) and the rings.
struct StreamRow {
  const unsigned char* psrc;     // &src[y][x]
  float* pimg;                   // &out_img[y - 2][x]
  float* pgx;                    // &out_gx[y - 5][x]
  float* pgy;
};

// One input row.  PH = (y - first row of the task) mod 7: slot of the rings written by this row.
// BORDER: the task touches an image border (warp uniform); interior tasks carry no zero-band tests.
// u8 word of row y for this lane (0 outside the image).  The loads of a whole 7-row group are
// issued one group ahead of their use: a row's load is the head of its dependency chain and a
// DRAM round trip per row would otherwise be exposed 70 times per task.
template <bool BORDER>
__device__ __forceinline__ unsigned l0_stream_load(StreamRow& p, int spitch, int H, int y, bool xin) {
  unsigned w = 0u;
  if (BORDER) { if (xin && y >= 0 && y < H) w = __ldg(reinterpret_cast<const unsigned*>(p.psrc)); }
  else w = __ldg(reinterpret_cast<const unsigned*>(p.psrc));
  p.psrc += spitch;
  return w;
}

template <bool EXACT, bool BORDER, int PH>
__device__ __forceinline__ void l0_stream_row(StreamRow& p, unsigned w, int opitch, int W, int H,
                                              const TapsF& ts, const TapsF& tg, const TapsF& td, int x, int y,
                                              int ys, int ye, bool own, float (&S)[7][4],
                                              float (&HD)[7][4], float (&HG)[7][4]) {
  using G = StreamGeo;
  constexpr int RS = G::RS, RG = G::RG;
  // ---- u8 row y: own 4 pixels + 2 on either side -----------------------------------------------
  const unsigned wl = __shfl_up_sync(0xffffffffu, w, 1), wr = __shfl_down_sync(0xffffffffu, w, 1);
  float f[8];                                   // pixels x-2 .. x+5
  f[0] = u8_to_float(wl, 2); f[1] = u8_to_float(wl, 3);
  f[2] = u8_to_float(w, 0);  f[3] = u8_to_float(w, 1); f[4] = u8_to_float(w, 2); f[5] = u8_to_float(w, 3);
  f[6] = u8_to_float(wr, 0); f[7] = u8_to_float(wr, 1);
  // ---- horizontal Gaussian -> ring slot PH --------------------------------------------------------
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float a = 0.0f;
#pragma unroll
    for (int m = 0; m < 2 * RS + 1; ++m) a = smac<EXACT>(a, f[c + m], ts.k[m]);
    if (BORDER) { const int xg = x + c; if (xg < RS || xg >= W - RS) a = 0.0f; }
    S[PH][c] = a;
  }
  // ---- vertical Gaussian: L0 row y-2 from hs rows y-4 .. y ----------------------------------------
  const int yl = y - RS;
  float L[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int m = 0; m < 2 * RS + 1; ++m) {
    const int s = (PH + 7 - 2 * RS + m) % 7;    // slot written at row y-4+m
    if (EXACT) {
#pragma unroll
      for (int c = 0; c < 4; ++c) L[c] = smac<true>(L[c], S[s][c], ts.k[m]);
    } else {
      sfma2(L[0], L[1], S[s][0], S[s][1], ts, m);
      sfma2(L[2], L[3], S[s][2], S[s][3], ts, m);
    }
  }
  if (BORDER) { if (yl < RS || yl >= H - RS) { L[0] = L[1] = L[2] = L[3] = 0.0f; } }
  if (own && yl >= ys && yl < ye) *reinterpret_cast<float4*>(p.pimg) = make_float4(L[0], L[1], L[2], L[3]);
  p.pimg += opitch;
  // ---- horizontal DoG / Gaussian of L0 row y-2 -> ring slot PH -------------------------------------
  float Lw[10];                                 // L0 at columns x-3 .. x+6
  Lw[0] = __shfl_up_sync(0xffffffffu, L[1], 1); Lw[1] = __shfl_up_sync(0xffffffffu, L[2], 1);
  Lw[2] = __shfl_up_sync(0xffffffffu, L[3], 1);
  Lw[3] = L[0]; Lw[4] = L[1]; Lw[5] = L[2]; Lw[6] = L[3];
  Lw[7] = __shfl_down_sync(0xffffffffu, L[0], 1); Lw[8] = __shfl_down_sync(0xffffffffu, L[1], 1);
  Lw[9] = __shfl_down_sync(0xffffffffu, L[2], 1);
  float hd[4] = {0.0f, 0.0f, 0.0f, 0.0f}, hg[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int m = 0; m < 2 * RG + 1; ++m) {
    if (EXACT) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (m != RG) hd[c] = smac<true>(hd[c], Lw[c + m], td.k[m]);        // centre tap is 0
        hg[c] = smac<true>(hg[c], Lw[c + m], tg.k[m]);
      }
    } else {
      if (m != RG) {
        sfma2(hd[0], hd[1], Lw[m], Lw[m + 1], td, m);
        sfma2(hd[2], hd[3], Lw[m + 2], Lw[m + 3], td, m);
      }
      sfma2(hg[0], hg[1], Lw[m], Lw[m + 1], tg, m);
      sfma2(hg[2], hg[3], Lw[m + 2], Lw[m + 3], tg, m);
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (BORDER) { const int xg = x + c; if (xg < RG || xg >= W - RG) { hd[c] = 0.0f; hg[c] = 0.0f; } }
    HD[PH][c] = hd[c]; HG[PH][c] = hg[c];
  }
  // ---- vertical passes: gradients of row y-5 from ring rows y-8 .. y-2 ------------------------------
  const int yg = y - RS - RG;
  float gx[4] = {0.0f, 0.0f, 0.0f, 0.0f}, gy[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int m = 0; m < 2 * RG + 1; ++m) {
    const int s = (PH + 1 + m) % 7;             // slot written at row y-6+m  (L0 row y-8+m)
    if (EXACT) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        gx[c] = smac<true>(gx[c], HD[s][c], tg.k[m]);
        if (m != RG) gy[c] = smac<true>(gy[c], HG[s][c], td.k[m]);
      }
    } else {
      sfma2(gx[0], gx[1], HD[s][0], HD[s][1], tg, m);
      sfma2(gx[2], gx[3], HD[s][2], HD[s][3], tg, m);
      if (m != RG) {
        sfma2(gy[0], gy[1], HG[s][0], HG[s][1], td, m);
        sfma2(gy[2], gy[3], HG[s][2], HG[s][3], td, m);
      }
    }
  }
  if (BORDER) {
    if (yg < RG || yg >= H - RG) { gx[0] = gx[1] = gx[2] = gx[3] = 0.0f; gy[0] = gy[1] = gy[2] = gy[3] = 0.0f; }
  }
  if (own && yg >= ys && yg < ye) {
    *reinterpret_cast<float4*>(p.pgx) = make_float4(gx[0], gx[1], gx[2], gx[3]);
    *reinterpret_cast<float4*>(p.pgy) = make_float4(gy[0], gy[1], gy[2], gy[3]);
  }
  p.pgx += opitch; p.pgy += opitch;
}

template <bool EXACT, bool BORDER>
__device__ __forceinline__ void l0_stream_task(const unsigned char* __restrict__ src, int spitch, int W, int H,
                                               const TapsF& ts, const TapsF& tg, const TapsF& td,
                                               float* __restrict__ out_img, float* __restrict__ out_gx,
                                               float* __restrict__ out_gy, int opitch, int x, int ys, int ye,
                                               bool own) {
  using G = StreamGeo;
  float S[7][4], HD[7][4], HG[7][4];
#pragma unroll
  for (int s = 0; s < 7; ++s)
#pragma unroll
    for (int c = 0; c < 4; ++c) { S[s][c] = 0.0f; HD[s][c] = 0.0f; HG[s][c] = 0.0f; }
  const int y0 = ys - G::LAT, y1 = ye + G::LAT;                            // input rows [y0, y1)
  const bool xin = x >= 0 && x < W;
  StreamRow p;                                   // (never dereferenced out of range: loads and stores are guarded)
  p.psrc = src + (long long)y0 * spitch + x;
  p.pimg = out_img + (long long)(y0 - G::RS) * opitch + x;
  p.pgx = out_gx + (long long)(y0 - G::LAT) * opitch + x;
  p.pgy = out_gy + (long long)(y0 - G::LAT) * opitch + x;
  unsigned wn[7];                                 // the next group's words, in flight
#pragma unroll
  for (int k = 0; k < 7; ++k) wn[k] = l0_stream_load<BORDER>(p, spitch, H, y0 + k, xin);
  for (int y = y0; y < y1; y += 7) {
    unsigned wc[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) wc[k] = wn[k];
    if (y + 7 < y1) {
#pragma unroll
      for (int k = 0; k < 7; ++k) wn[k] = l0_stream_load<BORDER>(p, spitch, H, y + 7 + k, xin);
    }
    l0_stream_row<EXACT, BORDER, 0>(p, wc[0], opitch, W, H, ts, tg, td, x, y + 0, ys, ye, own, S, HD, HG);
    l0_stream_row<EXACT, BORDER, 1>(p, wc[1], opitch, W, H, ts, tg, td, x, y + 1, ys, ye, own, S, HD, HG);
    l0_stream_row<EXACT, BORDER, 2>(p, wc[2], opitch, W, H, ts, tg, td, x, y + 2, ys, ye, own, S, HD, HG);
    l0_stream_row<EXACT, BORDER, 3>(p, wc[3], opitch, W, H, ts, tg, td, x, y + 3, ys, ye, own, S, HD, HG);
    l0_stream_row<EXACT, BORDER, 4>(p, wc[4], opitch, W, H, ts, tg, td, x, y + 4, ys, ye, own, S, HD, HG);
    l0_stream_row<EXACT, BORDER, 5>(p, wc[5], opitch, W, H, ts, tg, td, x, y + 5, ys, ye, own, S, HD, HG);
    l0_stream_row<EXACT, BORDER, 6>(p, wc[6], opitch, W, H, ts, tg, td, x, y + 6, ys, ye, own, S, HD, HG);
  }
}

// One warp per (strip, segment); HS output rows per segment (HS + 10 a multiple of 7 wastes nothing).
// Tasks [task0, ntasks) in strip-major order within a segment row (so that a launch over whole
// segment rows = whole image rows, as the banded upload needs).
template <bool EXACT>
__global__ void __launch_bounds__(128)
l0_stream_kernel(const unsigned char* __restrict__ src, int spitch, int W, int H, int nstrips, int HS,
                 int task0, int ntasks, TapsF ts, TapsF tg, TapsF td, float* __restrict__ out_img,
                 float* __restrict__ out_gx, float* __restrict__ out_gy, int opitch) {
  using G = StreamGeo;
  pdl_wait();                                   // frame / pyramid slot may still be in use by the previous kernel
  const int lane = threadIdx.x & 31;
  const int task = task0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (task >= ntasks) { pdl_launch_dependents(); return; }              // whole warps
  const int seg = task / nstrips, strip = task - seg * nstrips;
  const int xs = strip * G::OWN - G::HALO;                               // first column of the strip
  const int x = xs + G::C * lane;                                        // first column of this lane
  const int ys = seg * HS, ye = (ys + HS < H) ? ys + HS : H;
  const bool own = lane >= G::HALO / G::C && lane < 32 - G::HALO / G::C && x < W;
  // the last rows of the 7-row groups may run past ye + 5: keep them inside the image for the
  // unguarded loads of the interior variant
  const int rows7 = (ye - ys + 2 * G::LAT + 6) / 7 * 7;
  const bool border = xs < 8 || xs + G::SW + 8 > W || ys - G::LAT < 8 || ys - G::LAT + rows7 + 8 > H;
  if (border) l0_stream_task<EXACT, true>(src, spitch, W, H, ts, tg, td, out_img, out_gx, out_gy, opitch, x, ys, ye, own);
  else l0_stream_task<EXACT, false>(src, spitch, W, H, ts, tg, td, out_img, out_gx, out_gy, opitch, x, ys, ye, own);
  pdl_launch_dependents();
}
