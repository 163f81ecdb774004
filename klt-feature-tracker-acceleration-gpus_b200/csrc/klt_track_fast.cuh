// klt_track_fast.cuh -- production (FMA-mode) tracker kernel, included by klt_dev.cu.
//
// Same algorithm as track_kernel (reference src/V1/trackFeatures.c:381-486 per level and
// :1343-1437 across levels) re-mapped for latency: the warp-per-feature kernel is bound by a
// long dependent instruction chain (ncu: 2.6 k warp instructions per feature, 25 % issue
// utilisation, every first touch of a level a DRAM miss).  Here
//   * 8 lanes own one feature, 4 features share a warp: lane r of a group handles window row
//     r - WH/2 (WW pixels, all independent => deep memory-level parallelism);
//   * the bilinear weights are computed once per window position instead of once per pixel
//     (the fractional offsets of x+i and x are equal up to float rounding);
//   * the five window sums are reduced with 3 xor-shuffles over the 8 lanes, every lane ends
//     up with the totals and solves the 2x2 system redundantly (no broadcast);
//   * the footprint of the next finer level is prefetched into L2 while the current level
//     iterates.
// Windows up to 15 x 8*RPL rows are supported through the WW / RPL template parameters; the
// bit-exact mode and larger windows stay on track_kernel.
#pragma once

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ float group_sum8(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

// one window row of bilinear samples: out[i] = interp(img, xt+i+ax, yrow+ay), i = 0..WW-1
template <int WW>
__device__ __forceinline__ void sample_row(const float* __restrict__ img, int pitch, int xt, int yrow,
                                           float ax, float ay, float* out) {
  const float* p0 = img + (size_t)yrow * pitch + xt;
  const float* p1 = p0 + pitch;
  float a[WW + 1], b[WW + 1];
#pragma unroll
  for (int i = 0; i <= WW; ++i) { a[i] = __ldg(p0 + i); b[i] = __ldg(p1 + i); }
  const float w00 = (1.0f - ax) * (1.0f - ay), w01 = ax * (1.0f - ay), w10 = (1.0f - ax) * ay, w11 = ax * ay;
#pragma unroll
  for (int i = 0; i < WW; ++i)
    out[i] = fmaf(w11, b[i + 1], fmaf(w10, b[i], fmaf(w01, a[i + 1], w00 * a[i])));
}

// WW: window width (odd, <= 15); RPL: window rows per lane (window height <= 8 * RPL)
template <int WW, int RPL>
__global__ void __launch_bounds__(128)
track_fast_kernel(PyrView p1, PyrView p2, TrackArgs a, int n, FeatIO io,
                  unsigned long long* __restrict__ live_total) {
  const int lane = threadIdx.x & 31;
  const int r8 = lane & 7;                                   // lane within the feature group
  const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;   // feature of this group
  const int wh = a.wh, hw = WW / 2, hh = wh / 2;
  const float inv_npix = 0.0f;                               // (unused; residue uses division below)
  (void)inv_npix;

  bool alive = false;                                        // feature still being tracked
  float xloc = 0.0f, yloc = 0.0f;
  if (f < n) {
    const int v0 = io.val[(size_t)f * io.istride];
    alive = v0 >= 0;                                         // only features that are not lost (:1346)
    if (alive) { xloc = io.x[(size_t)f * io.istride]; yloc = io.y[(size_t)f * io.istride]; }
  }
  {
    const unsigned bal = __ballot_sync(0xffffffffu, alive && r8 == 0);
    if (lane == 0 && bal) atomicAdd(live_total, (unsigned long long)__popc(bal));
  }
  if (!__any_sync(0xffffffffu, alive)) return;

  for (int r = a.nlevels - 1; r >= 0; --r) { xloc = xloc / a.ss; yloc = yloc / a.ss; }
  float xout = xloc, yout = yloc;
  int status = KLT_TRACKED;
  bool running = alive;                                      // false once a level returned SMALL_DET / OOB

  // rows of the window this lane owns: j = r8 + 8*k - hh  (valid if r8 + 8*k < wh)
  for (int r = a.nlevels - 1; r >= 0; --r) {
    if (running) { xloc *= a.ss; yloc *= a.ss; xout *= a.ss; yout *= a.ss; }
    const int nc = p1.ncols[r], nr = p1.nrows[r], pitch = p1.pitch[r];
    const float* __restrict__ i1 = p1.img[r];
    const float* __restrict__ gx1 = p1.gx[r];
    const float* __restrict__ gy1 = p1.gy[r];
    const float* __restrict__ i2 = p2.img[r];
    const float* __restrict__ gx2 = p2.gx[r];
    const float* __restrict__ gy2 = p2.gy[r];

    // prefetch the footprint of the next finer level around the predicted positions
    if (running && r > 0) {
      const int pn = p1.pitch[r - 1], ncn = p1.ncols[r - 1], nrn = p1.nrows[r - 1];
      const int px1 = (int)(xloc * a.ss) - hw, py1 = (int)(yloc * a.ss) - hh + r8;
      const int px2 = (int)(xout * a.ss) - hw, py2 = (int)(yout * a.ss) - hh + r8;
      if (px1 >= 0 && py1 >= 0 && px1 + WW < ncn && py1 < nrn) {
        const size_t o = (size_t)py1 * pn + px1;
        prefetch_l2(p1.img[r - 1] + o); prefetch_l2(p1.img[r - 1] + o + WW);
        prefetch_l2(p1.gx[r - 1] + o);  prefetch_l2(p1.gx[r - 1] + o + WW);
        prefetch_l2(p1.gy[r - 1] + o);  prefetch_l2(p1.gy[r - 1] + o + WW);
      }
      if (px2 >= 0 && py2 >= 0 && px2 + WW < ncn && py2 < nrn) {
        const size_t o = (size_t)py2 * pn + px2;
        prefetch_l2(p2.img[r - 1] + o); prefetch_l2(p2.img[r - 1] + o + WW);
        prefetch_l2(p2.gx[r - 1] + o);  prefetch_l2(p2.gx[r - 1] + o + WW);
        prefetch_l2(p2.gy[r - 1] + o);  prefetch_l2(p2.gy[r - 1] + o + WW);
      }
    }

    // ---- _trackFeature at this level ----------------------------------------------------
    const float x1 = xloc, y1 = yloc;
    float x2 = xout, y2 = yout;
    int iteration = 0;
    float dx = 0.0f, dy = 0.0f;
    float t_i[RPL][WW], t_gx[RPL][WW], t_gy[RPL][WW];
    bool iterating = running;                                // this group still inside the do-while
    int lvl_status = KLT_TRACKED;

    if (iterating && window_oob(x1, y1, hw, hh, nc, nr)) { lvl_status = KLT_OOB; iterating = false; }
    if (iterating && window_oob(x2, y2, hw, hh, nc, nr)) { lvl_status = KLT_OOB; iterating = false; }
    if (iterating) {
      const int xt = (int)x1, yt = (int)y1;
      const float ax = x1 - (float)xt, ay = y1 - (float)yt;
#pragma unroll
      for (int k = 0; k < RPL; ++k) {
        const int row = r8 + 8 * k;
        if (row < wh) {
          sample_row<WW>(i1, pitch, xt - hw, yt - hh + row, ax, ay, t_i[k]);
          sample_row<WW>(gx1, pitch, xt - hw, yt - hh + row, ax, ay, t_gx[k]);
          sample_row<WW>(gy1, pitch, xt - hw, yt - hh + row, ax, ay, t_gy[k]);
        }
      }
    }

    while (__any_sync(0xffffffffu, iterating)) {
      float gxx = 0.0f, gxy = 0.0f, gyy = 0.0f, ex = 0.0f, ey = 0.0f;
      if (iterating) {
        const int xt = (int)x2, yt = (int)y2;
        const float ax = x2 - (float)xt, ay = y2 - (float)yt;
#pragma unroll
        for (int k = 0; k < RPL; ++k) {
          const int row = r8 + 8 * k;
          if (row < wh) {
            float s_i[WW], s_gx[WW], s_gy[WW];
            sample_row<WW>(i2, pitch, xt - hw, yt - hh + row, ax, ay, s_i);
            sample_row<WW>(gx2, pitch, xt - hw, yt - hh + row, ax, ay, s_gx);
            sample_row<WW>(gy2, pitch, xt - hw, yt - hh + row, ax, ay, s_gy);
#pragma unroll
            for (int i = 0; i < WW; ++i) {
              const float df = t_i[k][i] - s_i[i];
              const float sx = t_gx[k][i] + s_gx[i];
              const float sy = t_gy[k][i] + s_gy[i];
              gxx = fmaf(sx, sx, gxx); gxy = fmaf(sx, sy, gxy); gyy = fmaf(sy, sy, gyy);
              ex = fmaf(df, sx, ex); ey = fmaf(df, sy, ey);
            }
          }
        }
      }
      gxx = group_sum8(gxx); gxy = group_sum8(gxy); gyy = group_sum8(gyy);
      ex = group_sum8(ex); ey = group_sum8(ey);
      if (iterating) {
        ex *= a.step_factor; ey *= a.step_factor;
        const float det = gxx * gyy - gxy * gxy;
        if (det < a.min_determinant) {
          lvl_status = KLT_SMALL_DET; iterating = false;
        } else {
          dx = (gyy * ex - gxy * ey) / det;
          dy = (gxx * ey - gxy * ex) / det;
          x2 += dx; y2 += dy;
          ++iteration;
          const bool again = (fabsf(dx) >= a.min_displacement || fabsf(dy) >= a.min_displacement) &&
                             iteration < a.max_iterations;
          if (!again) iterating = false;
          else if (window_oob(x2, y2, hw, hh, nc, nr)) { lvl_status = KLT_OOB; iterating = false; }
        }
      }
    }

    // after the loop (:459-474): bounds of the final position, then the residue
    bool need_res = false;
    if (running) {
      if (window_oob(x2, y2, hw, hh, nc, nr)) lvl_status = KLT_OOB;
      need_res = (lvl_status == KLT_TRACKED);
    }
    if (__any_sync(0xffffffffu, need_res)) {
      float sum = 0.0f;
      if (need_res) {
        const int xt = (int)x2, yt = (int)y2;
        const float ax = x2 - (float)xt, ay = y2 - (float)yt;
#pragma unroll
        for (int k = 0; k < RPL; ++k) {
          const int row = r8 + 8 * k;
          if (row < wh) {
            float s_i[WW];
            sample_row<WW>(i2, pitch, xt - hw, yt - hh + row, ax, ay, s_i);
#pragma unroll
            for (int i = 0; i < WW; ++i) sum += fabsf(t_i[k][i] - s_i[i]);
          }
        }
      }
      sum = group_sum8(sum);
      if (need_res && sum / (float)(WW * wh) > a.max_residue) lvl_status = KLT_LARGE_RESIDUE;
    }
    if (running) {
      int v;                                               // return value of _trackFeature (:479-484)
      if (lvl_status == KLT_SMALL_DET) v = KLT_SMALL_DET;
      else if (lvl_status == KLT_OOB) v = KLT_OOB;
      else if (lvl_status == KLT_LARGE_RESIDUE) v = KLT_LARGE_RESIDUE;
      else if (iteration >= a.max_iterations) v = KLT_MAX_ITERATIONS;
      else v = KLT_TRACKED;
      status = v;
      xout = x2; yout = y2;
      if (v == KLT_SMALL_DET || v == KLT_OOB) running = false;       // :1378
    }
    if (!__any_sync(0xffffffffu, running)) break;
  }

  if (alive && r8 == 0) {                                   // record (:1383-1437)
    const bool outside = (xout < (float)a.borderx || xout > (float)(a.ncols - 1 - a.borderx) ||
                          yout < (float)a.bordery || yout > (float)(a.nrows - 1 - a.bordery));
    const size_t o = (size_t)f * io.ostride;
    if (status == KLT_OOB || outside) { io.ox[o] = -1.0f; io.oy[o] = -1.0f; io.oval[o] = KLT_OOB; }
    else if (status != KLT_TRACKED) { io.ox[o] = -1.0f; io.oy[o] = -1.0f; io.oval[o] = status; }
    else { io.ox[o] = xout; io.oy[o] = yout; io.oval[o] = KLT_TRACKED; }
  }
}

// ---------------------------------------------------------------------------------------------
// 7x7 window (the default and the window of every BASELINE config): vectorised footprint loads.
//
// ncu on track_fast_kernel: 14.2 M L1 sectors for 0.52 M load requests -- each scalar load of a
// warp touches 32 different image rows, so the kernel is bound by L1TEX wavefronts, not by DRAM
// or arithmetic.  A 7x7 window at a fractional position reads an 8x8 pixel footprint: 8 rows
// for the 8 lanes of a feature group.  Lane r loads footprint row r as three aligned float4
// (the 8 pixels straddle at most three 16 B chunks), realigns them in registers with two levels
// of selects (offset 0..3), interpolates horizontally, and obtains the row below from lane r+1
// with a shuffle for the vertical interpolation.  Loads per lane and image: 3 x LDG.128 instead
// of 16 x LDG.32.
// ---------------------------------------------------------------------------------------------
struct Foot7 {
  int off;        // element offset of the first aligned chunk of this lane's row
  int o;          // 0..3: position of footprint column 0 inside the first chunk
  float ax, ay;
};
__device__ __forceinline__ Foot7 foot7_setup(float x, float y, int pitch, int r8) {
  const int xt = (int)x, yt = (int)y;
  Foot7 f;
  const int xs = xt - 3;                         // >= 0 (bounds were checked)
  f.o = xs & 3;
  f.off = (yt - 3 + r8) * pitch + (xs & ~3);
  f.ax = x - (float)xt;
  f.ay = y - (float)yt;
  return f;
}
// raw footprint row of this lane: three aligned 16 B chunks
struct Row12 { float4 c0, c1, c2; };
__device__ __forceinline__ Row12 foot7_load(const float* __restrict__ img, int off) {
  const float4* p = reinterpret_cast<const float4*>(img + off);
  Row12 r;
  r.c0 = __ldg(p); r.c1 = __ldg(p + 1); r.c2 = __ldg(p + 2);
  return r;
}
// window row r8 of the bilinear samples (valid for r8 < 7) from the raw row of this lane and,
// through a shuffle, the row below.  gmask: the 8 lanes of this feature group (whole groups call
// this together).
__device__ __forceinline__ void foot7_interp(const Row12& r, const Foot7& f, unsigned gmask, float* out) {
  const float v[12] = {r.c0.x, r.c0.y, r.c0.z, r.c0.w, r.c1.x, r.c1.y, r.c1.z, r.c1.w,
                       r.c2.x, r.c2.y, r.c2.z, r.c2.w};
  float t[10], q[8];
  const bool s2 = (f.o & 2) != 0, s1 = (f.o & 1) != 0;
#pragma unroll
  for (int j = 0; j < 10; ++j) t[j] = s2 ? v[j + 2] : v[j];
#pragma unroll
  for (int j = 0; j < 8; ++j) q[j] = s1 ? t[j + 1] : t[j];
  float h[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) h[i] = fmaf(f.ax, q[i + 1] - q[i], q[i]);
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const float below = __shfl_down_sync(gmask, h[i], 1, 8);
    out[i] = fmaf(f.ay, below - h[i], h[i]);
  }
}
__device__ __forceinline__ void foot7_sample(const float* __restrict__ img, const Foot7& f, unsigned gmask,
                                             float* out) {
  const Row12 r = foot7_load(img, f.off);
  foot7_interp(r, f, gmask, out);
}

// L2 prefetch of this lane's footprint row of one level, both frames, around (x, y): the three
// 32 B sectors around the aligned 48 B the loads will touch if the feature moves by a few pixels.
__device__ __forceinline__ void foot7_prefetch(const PyrView& p1, const PyrView& p2, int r, float x, float y,
                                               int r8) {
  const int pitch = p1.pitch[r];
  const int xs = (((int)x) - 3) & ~3, row = ((int)y) - 3 + r8;
  if (xs < 4 || row < 0 || row >= p1.nrows[r] || xs + 16 > pitch) return;
  const size_t o = (size_t)row * pitch + xs;
#pragma unroll
  for (int k = -4; k <= 12; k += 8) {
    prefetch_l2(p1.img[r] + o + k); prefetch_l2(p1.gx[r] + o + k); prefetch_l2(p1.gy[r] + o + k);
    prefetch_l2(p2.img[r] + o + k); prefetch_l2(p2.gx[r] + o + k); prefetch_l2(p2.gy[r] + o + k);
  }
}

// FPW: features per warp (1, 2 or 4).  Fewer features per warp = more warps in flight for the
// same work: the kernel is latency bound (issue slots are ~10 % used), so idle lanes are free.
template <int FPW>
__global__ void __launch_bounds__(128)
track7_kernel(PyrView p1, PyrView p2, TrackArgs a, int n, FeatIO io,
              unsigned long long* __restrict__ live_total) {
  constexpr int WW = 7, hw = 3, hh = 3;
  const int lane = threadIdx.x & 31;
  const int r8 = lane & 7;
  const bool rowlane = r8 < 7;                                // lane 7 only feeds the shuffle
  const unsigned gmask = 0xFFu << (lane & 24);                // the 8 lanes of this feature group
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int grp = lane >> 3;
  const int f = grp < FPW ? warp_global * FPW + grp : n;      // groups >= FPW stay idle

  pdl_wait();                                                 // pyramids and features come from earlier kernels
  bool alive = false;
  float xloc = 0.0f, yloc = 0.0f;
  if (f < n) {
    alive = io.val[(size_t)f * io.istride] >= 0;              // only features that are not lost (:1346)
    if (a.pass == 2 && a.done[f]) alive = false;              // answered by the early pass
    if (a.pass == 1 && !alive && r8 == 0) a.done[f] = 1;      // nothing to do for a lost feature
    if (alive) { xloc = io.x[(size_t)f * io.istride]; yloc = io.y[(size_t)f * io.istride]; }
  }
  if (a.pass != 2) {
    const unsigned bal = __ballot_sync(0xffffffffu, alive && r8 == 0);
    if (lane == 0 && bal) atomicAdd(live_total, (unsigned long long)__popc(bal));
  }
  bool deferred = false;                                      // pass 1: footprint beyond the rows that exist
  if (!__any_sync(0xffffffffu, alive)) return;

  // Every first touch of a level is a DRAM miss (two 132 MB pyramid sets do not fit in L2) and the
  // levels are a dependent chain: pull the footprints of all the finer levels towards L2 now, while
  // the coarsest level is being worked on.  The feature moves by a few pixels at most.
  if (alive && a.prefetch) {
    float px = xloc, py = yloc;
    for (int r = 0; r < a.nlevels - 1; ++r) {
      foot7_prefetch(p1, p2, r, px, py, r8);
      px = px / a.ss; py = py / a.ss;
    }
  }
  for (int r = a.nlevels - 1; r >= 0; --r) { xloc = xloc / a.ss; yloc = yloc / a.ss; }
  float xout = xloc, yout = yloc;
  int status = KLT_TRACKED;
  bool running = alive;

  for (int r = a.nlevels - 1; r >= 0; --r) {
    if (running) { xloc *= a.ss; yloc *= a.ss; xout *= a.ss; yout *= a.ss; }
    const int nc = p1.ncols[r], nr = p1.nrows[r], pitch = p1.pitch[r];
    const float* __restrict__ i1 = p1.img[r];
    const float* __restrict__ gx1 = p1.gx[r];
    const float* __restrict__ gy1 = p1.gy[r];
    const float* __restrict__ i2 = p2.img[r];
    const float* __restrict__ gx2 = p2.gx[r];
    const float* __restrict__ gy2 = p2.gy[r];

    const float x1 = xloc, y1 = yloc;
    float x2 = xout, y2 = yout;
    int iteration = 0;
    float dx = 0.0f, dy = 0.0f;
    float t_i[WW], t_gx[WW], t_gy[WW];
    bool iterating = running;
    int lvl_status = KLT_TRACKED;

    if (iterating && window_oob(x1, y1, hw, hh, nc, nr)) { lvl_status = KLT_OOB; iterating = false; }
    if (iterating && window_oob(x2, y2, hw, hh, nc, nr)) { lvl_status = KLT_OOB; iterating = false; }
    // pass 1: the footprint rows (int)y2 - 3 .. (int)y2 + 4 of the new frame must already exist
    const int row_lim = a.pass == 1 ? a.row_limit[r] : 0x7fffffff;
    if (iterating && (int)y2 + 4 >= row_lim) { deferred = true; iterating = false; running = false; }
    // raw footprint rows of frame 2, kept across iterations: sub-pixel Newton updates and the
    // residue pass usually stay on the same 8x8 integer footprint, so nothing is re-read
    Row12 c_i, c_gx, c_gy;
    int c_off = -1;
    // the sampling routines shuffle inside the 8-lane group, so whole groups enter them together
    // (iterating is uniform within a group)
    if (iterating) {
      // all 18 loads of the level (template + first frame-2 footprint) are issued before the first
      // use, so the level pays one memory round trip instead of two
      const Foot7 ft = foot7_setup(x1, y1, pitch, r8);
      const Foot7 f2 = foot7_setup(x2, y2, pitch, r8);
      const Row12 r_i = foot7_load(i1, ft.off), r_gx = foot7_load(gx1, ft.off), r_gy = foot7_load(gy1, ft.off);
      c_i = foot7_load(i2, f2.off); c_gx = foot7_load(gx2, f2.off); c_gy = foot7_load(gy2, f2.off);
      c_off = f2.off;
      foot7_interp(r_i, ft, gmask, t_i);
      foot7_interp(r_gx, ft, gmask, t_gx);
      foot7_interp(r_gy, ft, gmask, t_gy);
    }

    while (__any_sync(0xffffffffu, iterating)) {
      float gxx = 0.0f, gxy = 0.0f, gyy = 0.0f, ex = 0.0f, ey = 0.0f;
      if (iterating) {
        const Foot7 ft = foot7_setup(x2, y2, pitch, r8);
        if (ft.off != c_off) {                               // uniform within the group
          if ((int)y2 + 4 >= row_lim) { deferred = true; iterating = false; running = false; }
          else {
            c_i = foot7_load(i2, ft.off); c_gx = foot7_load(gx2, ft.off); c_gy = foot7_load(gy2, ft.off);
            c_off = ft.off;
          }
        }
      }
      if (iterating) {
        const Foot7 ft = foot7_setup(x2, y2, pitch, r8);
        float s_i[WW], s_gx[WW], s_gy[WW];
        foot7_interp(c_i, ft, gmask, s_i);
        foot7_interp(c_gx, ft, gmask, s_gx);
        foot7_interp(c_gy, ft, gmask, s_gy);
        if (rowlane) {
#pragma unroll
          for (int i = 0; i < WW; ++i) {
            const float df = t_i[i] - s_i[i];
            const float sx = t_gx[i] + s_gx[i];
            const float sy = t_gy[i] + s_gy[i];
            gxx = fmaf(sx, sx, gxx); gxy = fmaf(sx, sy, gxy); gyy = fmaf(sy, sy, gyy);
            ex = fmaf(df, sx, ex); ey = fmaf(df, sy, ey);
          }
        }
      }
      gxx = group_sum8(gxx); gxy = group_sum8(gxy); gyy = group_sum8(gyy);
      ex = group_sum8(ex); ey = group_sum8(ey);
      if (iterating) {
        ex *= a.step_factor; ey *= a.step_factor;
        const float det = gxx * gyy - gxy * gxy;
        if (det < a.min_determinant) {
          lvl_status = KLT_SMALL_DET; iterating = false;
        } else {
          const float inv = __frcp_rn(det);                   // one reciprocal instead of two divisions
          dx = (gyy * ex - gxy * ey) * inv;
          dy = (gxx * ey - gxy * ex) * inv;
          x2 += dx; y2 += dy;
          ++iteration;
          const bool again = (fabsf(dx) >= a.min_displacement || fabsf(dy) >= a.min_displacement) &&
                             iteration < a.max_iterations;
          if (!again) iterating = false;
          else if (window_oob(x2, y2, hw, hh, nc, nr)) { lvl_status = KLT_OOB; iterating = false; }
        }
      }
    }

    bool need_res = false;
    if (running) {
      if (window_oob(x2, y2, hw, hh, nc, nr)) lvl_status = KLT_OOB;
      need_res = (lvl_status == KLT_TRACKED);
      if (need_res && foot7_setup(x2, y2, pitch, r8).off != c_off && (int)y2 + 4 >= row_lim) {
        deferred = true; running = false; need_res = false;   // the residue footprint does not exist yet
      }
    }
    if (__any_sync(0xffffffffu, need_res)) {
      float sum = 0.0f;
      if (need_res) {
        const Foot7 ft = foot7_setup(x2, y2, pitch, r8);
        if (ft.off != c_off) { c_i = foot7_load(i2, ft.off); c_off = ft.off; }
        float s_i[WW];
        foot7_interp(c_i, ft, gmask, s_i);
        if (rowlane) {
#pragma unroll
          for (int i = 0; i < WW; ++i) sum += fabsf(t_i[i] - s_i[i]);
        }
      }
      sum = group_sum8(sum);
      if (need_res && sum * (1.0f / 49.0f) > a.max_residue) lvl_status = KLT_LARGE_RESIDUE;
    }
    if (running) {
      int v;
      if (lvl_status == KLT_SMALL_DET) v = KLT_SMALL_DET;
      else if (lvl_status == KLT_OOB) v = KLT_OOB;
      else if (lvl_status == KLT_LARGE_RESIDUE) v = KLT_LARGE_RESIDUE;
      else if (iteration >= a.max_iterations) v = KLT_MAX_ITERATIONS;
      else v = KLT_TRACKED;
      status = v;
      xout = x2; yout = y2;
      if (v == KLT_SMALL_DET || v == KLT_OOB) running = false;
    }
    if (!__any_sync(0xffffffffu, running)) break;
  }

  if (alive && r8 == 0 && a.pass == 1) a.done[f] = deferred ? 0 : 1;
  if (alive && r8 == 0 && !deferred) {
    const bool outside = (xout < (float)a.borderx || xout > (float)(a.ncols - 1 - a.borderx) ||
                          yout < (float)a.bordery || yout > (float)(a.nrows - 1 - a.bordery));
    const size_t o = (size_t)f * io.ostride;
    if (status == KLT_OOB || outside) { io.ox[o] = -1.0f; io.oy[o] = -1.0f; io.oval[o] = KLT_OOB; }
    else if (status != KLT_TRACKED) { io.ox[o] = -1.0f; io.oy[o] = -1.0f; io.oval[o] = status; }
    else { io.ox[o] = xout; io.oy[o] = yout; io.oval[o] = KLT_TRACKED; }
  }
}

// ---------------------------------------------------------------------------------------------
// 7x7 window, one WARP per feature (track7w_kernel): the latency-optimised tracker.
//
// track7_kernel (8 lanes per feature) issues ~350 dependent-ish instructions per Newton iteration
// from each warp at 1.7 warps per scheduler: measured, a pass over half of the features takes as
// long as a pass over all of them (27 us) -- the kernel is a latency chain, not a throughput
// problem, and the GPU is idle otherwise.  Here the 8x8 pixel footprint is spread over all 32
// lanes (lane = 4 * row + column pair): each lane holds the three pixels it needs per image,
// interpolates two window columns (10 instructions per image instead of ~53), the five sums are
// reduced with 5 xor-shuffles, every lane solves the 2x2 system.  ~135 instructions per iteration
// and 4x as many warps in flight.  Same per-sample arithmetic as track7_kernel; the sums are
// added in a different (tree) order.
// ---------------------------------------------------------------------------------------------
struct Pix3 { float a, b, c; };
__device__ __forceinline__ Pix3 pix3_load(const float* __restrict__ img, int off) {
  const float* p = img + off;
  Pix3 r;
  r.a = __ldg(p); r.b = __ldg(p + 1); r.c = __ldg(p + 2);
  return r;
}
// L2 eviction policies for the gathers.  The footprints read from the NEW frame's pyramids are
// (up to the integer cell the feature ends in) the footprints the next frame's call reads from its
// PREVIOUS pyramids; marked evict_last they survive the ~140 MB the next pyramid build streams
// through the 126 MB L2, so half of the next call's gathers hit L2 instead of DRAM.  The previous
// frame's footprints are dead after this call: evict_first.
__device__ __forceinline__ unsigned long long l2_policy(int kind) {   // 0 normal, 1 evict_last, 2 evict_first
  unsigned long long p;
  if (kind == 1) asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else if (kind == 2) asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  else asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float ldg_policy(const float* p, unsigned long long pol) {
  float v;
  asm("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}
// two loads per lane; the third pixel (column 2c + 2) is the first pixel of the lane to the right.
// Lane c = 3 receives the next row's first pixel instead, which it never uses (its second window
// column does not exist).  A third of the load instructions (and of their L1 wavefronts: every load
// instruction of the warp touches 8 rows = 8+ cache lines) for one shuffle.  Warp-uniform callers only.
__device__ __forceinline__ Pix3 pix3_load(const float* __restrict__ img, int off, unsigned long long pol) {
  const float* p = img + off;
  Pix3 r;
  if ((off & 1) == 0) {        // footprint starts at an even column (warp uniform: row pitches are even): one 64-bit load
    asm("ld.global.nc.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(r.a), "=f"(r.b) : "l"(p), "l"(pol));
  } else {
    r.a = ldg_policy(p, pol); r.b = ldg_policy(p + 1, pol);
  }
  r.c = __shfl_down_sync(0xffffffffu, r.a, 1);
  return r;
}
// bilinear samples of this lane's two window columns in its window row; the row below comes from
// lane + 4 (lanes 28..31 hold footprint row 7 and only feed the shuffle)
__device__ __forceinline__ void pix3_interp(const Pix3& p, float ax, float ay, float& o0, float& o1) {
  const float h0 = fmaf(ax, p.b - p.a, p.a), h1 = fmaf(ax, p.c - p.b, p.b);
  const float b0 = __shfl_down_sync(0xffffffffu, h0, 4), b1 = __shfl_down_sync(0xffffffffu, h1, 4);
  o0 = fmaf(ay, b0 - h0, h0);
  o1 = fmaf(ay, b1 - h1, h1);
}
__device__ __forceinline__ float warp_sum32(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// totals of four per-lane values in every lane with a splitting butterfly: the xor-16 exchange
// halves the values a lane carries from four to two, the xor-8 exchange to one, three plain steps
// finish it and four broadcasts hand every total to every lane: 10 shuffles, 6 adds, 6 selects
// instead of 20 shuffles and 20 adds.
__device__ __forceinline__ void warp_sum4x(float& g0, float& g1, float& g2, float& g3, int lane) {
  const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
  const float r0 = __shfl_xor_sync(0xffffffffu, hi16 ? g0 : g2, 16);
  const float r1 = __shfl_xor_sync(0xffffffffu, hi16 ? g1 : g3, 16);
  const float k0 = (hi16 ? g2 : g0) + r0, k1 = (hi16 ? g3 : g1) + r1;    // lower half: g0, g1; upper half: g2, g3
  float k = (hi8 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, hi8 ? k0 : k1, 8);
  k += __shfl_xor_sync(0xffffffffu, k, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);           // lanes 0-7: g0, 8-15: g1, 16-23: g2, 24-31: g3
  g0 = __shfl_sync(0xffffffffu, k, 0); g1 = __shfl_sync(0xffffffffu, k, 8);
  g2 = __shfl_sync(0xffffffffu, k, 16); g3 = __shfl_sync(0xffffffffu, k, 24);
}

__global__ void __launch_bounds__(128)
track7w_kernel(PyrView p1, PyrView p2, TrackArgs a, int n, FeatIO io,
               unsigned long long* __restrict__ live_total) {
  constexpr int hw = 3, hh = 3;
  const int lane = threadIdx.x & 31;
  const int r = lane >> 2, c = lane & 3;                     // footprint row, column pair
  const bool v0 = r < 7, v1 = r < 7 && c < 3;                // window samples (2c, r) and (2c + 1, r)
  const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned long long pol1 = l2_policy(a.l2_keep ? 2 : 0), pol2 = l2_policy(a.l2_keep ? 1 : 0);
  pdl_wait();                                                 // pyramids and features come from earlier kernels
  if (f >= n) return;
  if (io.val[(size_t)f * io.istride] < 0) return;             // only features that are not lost (:1346)
  if (lane == 0) atomicAdd(live_total, 1ULL);
  float xloc = io.x[(size_t)f * io.istride], yloc = io.y[(size_t)f * io.istride];
  for (int l = a.nlevels - 1; l >= 0; --l) { xloc = xloc / a.ss; yloc = yloc / a.ss; }
  float xout = xloc, yout = yloc;
  int status = KLT_TRACKED;

  for (int l = a.nlevels - 1; l >= 0; --l) {
    xloc *= a.ss; yloc *= a.ss; xout *= a.ss; yout *= a.ss;
    const int nc = p1.ncols[l], nr = p1.nrows[l], pitch = p1.pitch[l];
    const float* __restrict__ i2 = p2.img[l];
    const float* __restrict__ gx2 = p2.gx[l];
    const float* __restrict__ gy2 = p2.gy[l];
    const float x1 = xloc, y1 = yloc;
    float x2 = xout, y2 = yout;
    int iteration = 0, lvl_status = KLT_TRACKED;
    bool iterating = true;
    if (window_oob(x1, y1, hw, hh, nc, nr) || window_oob(x2, y2, hw, hh, nc, nr)) { lvl_status = KLT_OOB; iterating = false; }

    float t_i0 = 0.f, t_i1 = 0.f, t_gx0 = 0.f, t_gx1 = 0.f, t_gy0 = 0.f, t_gy1 = 0.f;
    Pix3 c_i, c_gx, c_gy;
    c_i.a = c_i.b = c_i.c = 0.f; c_gx = c_i; c_gy = c_i;
    int c_xt = -1000000, c_yt = -1000000;
    if (iterating) {
      // all 18 loads of the level are issued before the first use
      const int xt1 = (int)x1, yt1 = (int)y1, xt2 = (int)x2, yt2 = (int)y2;
      const int o1 = (yt1 - 3 + r) * pitch + (xt1 - 3 + 2 * c), o2 = (yt2 - 3 + r) * pitch + (xt2 - 3 + 2 * c);
      const Pix3 r_i = pix3_load(p1.img[l], o1, pol1), r_gx = pix3_load(p1.gx[l], o1, pol1), r_gy = pix3_load(p1.gy[l], o1, pol1);
      c_i = pix3_load(i2, o2, pol2); c_gx = pix3_load(gx2, o2, pol2); c_gy = pix3_load(gy2, o2, pol2);
      c_xt = xt2; c_yt = yt2;
      const float ax = x1 - (float)xt1, ay = y1 - (float)yt1;
      pix3_interp(r_i, ax, ay, t_i0, t_i1);
      pix3_interp(r_gx, ax, ay, t_gx0, t_gx1);
      pix3_interp(r_gy, ax, ay, t_gy0, t_gy1);
    }

    float dx = 0.0f, dy = 0.0f;
    while (iterating) {                                       // warp uniform
      const int xt = (int)x2, yt = (int)y2;
      if (xt != c_xt || yt != c_yt) {                         // the integer footprint moved: re-read it
        const int o2 = (yt - 3 + r) * pitch + (xt - 3 + 2 * c);
        c_i = pix3_load(i2, o2, pol2); c_gx = pix3_load(gx2, o2, pol2); c_gy = pix3_load(gy2, o2, pol2);
        c_xt = xt; c_yt = yt;
      }
      const float ax = x2 - (float)xt, ay = y2 - (float)yt;
      float s_i0, s_i1, s_gx0, s_gx1, s_gy0, s_gy1;
      pix3_interp(c_i, ax, ay, s_i0, s_i1);
      pix3_interp(c_gx, ax, ay, s_gx0, s_gx1);
      pix3_interp(c_gy, ax, ay, s_gy0, s_gy1);
      float gxx = 0.0f, gxy = 0.0f, gyy = 0.0f, ex = 0.0f, ey = 0.0f;
      {
        const float df = v0 ? t_i0 - s_i0 : 0.0f, sx = v0 ? t_gx0 + s_gx0 : 0.0f, sy = v0 ? t_gy0 + s_gy0 : 0.0f;
        gxx = sx * sx; gxy = sx * sy; gyy = sy * sy; ex = df * sx; ey = df * sy;
      }
      {
        const float df = v1 ? t_i1 - s_i1 : 0.0f, sx = v1 ? t_gx1 + s_gx1 : 0.0f, sy = v1 ? t_gy1 + s_gy1 : 0.0f;
        gxx = fmaf(sx, sx, gxx); gxy = fmaf(sx, sy, gxy); gyy = fmaf(sy, sy, gyy);
        ex = fmaf(df, sx, ex); ey = fmaf(df, sy, ey);
      }
      warp_sum4x(gxx, gxy, gyy, ex, lane);
      ey = warp_sum32(ey);
      ex *= a.step_factor; ey *= a.step_factor;
      const float det = gxx * gyy - gxy * gxy;
      if (det < a.min_determinant) { lvl_status = KLT_SMALL_DET; break; }
      const float inv = __frcp_rn(det);
      dx = (gyy * ex - gxy * ey) * inv;
      dy = (gxx * ey - gxy * ex) * inv;
      x2 += dx; y2 += dy;
      ++iteration;
      const bool again = (fabsf(dx) >= a.min_displacement || fabsf(dy) >= a.min_displacement) &&
                         iteration < a.max_iterations;
      if (!again) break;
      if (window_oob(x2, y2, hw, hh, nc, nr)) { lvl_status = KLT_OOB; break; }
    }

    // after the loop (:459-474): bounds of the final position, then the residue
    if (lvl_status == KLT_TRACKED && window_oob(x2, y2, hw, hh, nc, nr)) lvl_status = KLT_OOB;
    if (lvl_status == KLT_TRACKED) {
      const int xt = (int)x2, yt = (int)y2;
      if (xt != c_xt || yt != c_yt) c_i = pix3_load(i2, (yt - 3 + r) * pitch + (xt - 3 + 2 * c), pol2);
      float s_i0, s_i1;
      pix3_interp(c_i, x2 - (float)xt, y2 - (float)yt, s_i0, s_i1);
      float sum = (v0 ? fabsf(t_i0 - s_i0) : 0.0f) + (v1 ? fabsf(t_i1 - s_i1) : 0.0f);
      sum = warp_sum32(sum);
      if (sum * (1.0f / 49.0f) > a.max_residue) lvl_status = KLT_LARGE_RESIDUE;
    }
    int v;                                                   // return value of _trackFeature (:479-484)
    if (lvl_status == KLT_SMALL_DET) v = KLT_SMALL_DET;
    else if (lvl_status == KLT_OOB) v = KLT_OOB;
    else if (lvl_status == KLT_LARGE_RESIDUE) v = KLT_LARGE_RESIDUE;
    else if (iteration >= a.max_iterations) v = KLT_MAX_ITERATIONS;
    else v = KLT_TRACKED;
    status = v;
    xout = x2; yout = y2;
    if (v == KLT_SMALL_DET || v == KLT_OOB) break;           // :1378
  }

  if (lane == 0) {                                           // record (:1383-1437)
    const bool outside = (xout < (float)a.borderx || xout > (float)(a.ncols - 1 - a.borderx) ||
                          yout < (float)a.bordery || yout > (float)(a.nrows - 1 - a.bordery));
    const size_t o = (size_t)f * io.ostride;
    if (status == KLT_OOB || outside) { io.ox[o] = -1.0f; io.oy[o] = -1.0f; io.oval[o] = KLT_OOB; }
    else if (status != KLT_TRACKED) { io.ox[o] = -1.0f; io.oy[o] = -1.0f; io.oval[o] = status; }
    else { io.ox[o] = xout; io.oy[o] = yout; io.oval[o] = KLT_TRACKED; }
  }
}

// ---------------------------------------------------------------------------------------------
// 7x7 window, one warp per feature, ONE 128-bit load per lane and image (track7v_kernel).
//
// ncu on track7w_kernel: every scalar load instruction of the warp touches the 8 footprint rows,
// i.e. 8+ cache lines = 8+ L1 wavefronts, and a level start issues 18 of them; dropping one load in
// three (the third pixel by shuffle) took the kernel from 27.1 to 24.7 us -- it is bound by the
// L1 wavefronts of its gathers, not by DRAM bytes (an L2 set-aside for the footprints changed the
// DRAM traffic but not the time) and not by its iteration count.  Here lane = 4 * row + q and lanes
// q = 0..2 load the three ALIGNED float4 chunks that cover the row's footprint (floats B .. B+11,
// B = (xt - 3) & ~3; the footprint starts at float m = (xt - 3) & 3 of them): one load instruction
// per image instead of two.  A lane then works on the window columns that start in its chunk,
// w = 4q + t - m for t = 0..3 (those in 0..6 are real), with the first float of the chunk to its
// right by shuffle.  The previous frame's samples are aligned to ITS m; they are shifted to the new
// frame's alignment by shuffles whenever that changes.  Per-sample arithmetic as in track7w_kernel;
// the sums run over up to four samples per lane before the xor-shuffle tree.
// ---------------------------------------------------------------------------------------------
struct Row5 { float x, y, z, w, n; };
__device__ __forceinline__ Row5 row5_load(const float* __restrict__ img, int off, bool active, unsigned long long pol) {
  Row5 r;
  r.x = r.y = r.z = r.w = 0.0f;
  if (active)
    asm("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(img + off), "l"(pol));
  r.n = __shfl_down_sync(0xffffffffu, r.x, 1);
  return r;
}
__device__ __forceinline__ void row5_interp(const Row5& p, float ax, float ay, float (&o)[4]) {
  const float h0 = fmaf(ax, p.y - p.x, p.x), h1 = fmaf(ax, p.z - p.y, p.y);
  const float h2 = fmaf(ax, p.w - p.z, p.z), h3 = fmaf(ax, p.n - p.w, p.w);
  const float b0 = __shfl_down_sync(0xffffffffu, h0, 4), b1 = __shfl_down_sync(0xffffffffu, h1, 4);
  const float b2 = __shfl_down_sync(0xffffffffu, h2, 4), b3 = __shfl_down_sync(0xffffffffu, h3, 4);
  o[0] = fmaf(ay, b0 - h0, h0); o[1] = fmaf(ay, b1 - h1, h1);
  o[2] = fmaf(ay, b2 - h2, h2); o[3] = fmaf(ay, b3 - h3, h3);
}
// out[t] = the sample that sits `delta` slots to the left in the row (slot = 4 * q + t); delta is
// warp uniform, -3 .. 3
__device__ __forceinline__ void row_shift(const float (&in)[4], int delta, int lane, float (&out)[4]) {
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int s = t - delta;                       // -3 .. 6
    const int t1 = s & 3, dq = (s - t1) >> 2;      // source slot within its lane, lane offset -1 / 0 / 1
    const float v = t1 == 0 ? in[0] : t1 == 1 ? in[1] : t1 == 2 ? in[2] : in[3];
    out[t] = __shfl_sync(0xffffffffu, v, (lane + dq) & 31);
  }
}

__global__ void __launch_bounds__(128)
track7v_kernel(PyrView p1, PyrView p2, TrackArgs a, int n, FeatIO io,
               unsigned long long* __restrict__ live_total) {
  constexpr int hw = 3, hh = 3;
  const int lane = threadIdx.x & 31;
  const int r = lane >> 2, q = lane & 3;                     // footprint row, aligned chunk of the row
  const bool ld = q < 3;
  const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned long long pol1 = l2_policy(a.l2_keep ? 2 : 0), pol2 = l2_policy(a.l2_keep ? 1 : 0);
  pdl_wait();                                                 // pyramids and features come from earlier kernels
  if (f >= n) return;
  if (io.val[(size_t)f * io.istride] < 0) return;             // only features that are not lost (:1346)
  if (lane == 0) atomicAdd(live_total, 1ULL);
  float xloc = io.x[(size_t)f * io.istride], yloc = io.y[(size_t)f * io.istride];
  for (int l = a.nlevels - 1; l >= 0; --l) { xloc = xloc / a.ss; yloc = yloc / a.ss; }
  float xout = xloc, yout = yloc;
  int status = KLT_TRACKED;

  for (int l = a.nlevels - 1; l >= 0; --l) {
    xloc *= a.ss; yloc *= a.ss; xout *= a.ss; yout *= a.ss;
    const int nc = p1.ncols[l], nr = p1.nrows[l], pitch = p1.pitch[l];
    const float* __restrict__ i2 = p2.img[l];
    const float* __restrict__ gx2 = p2.gx[l];
    const float* __restrict__ gy2 = p2.gy[l];
    const float x1 = xloc, y1 = yloc;
    float x2 = xout, y2 = yout;
    int iteration = 0, lvl_status = KLT_TRACKED;
    bool iterating = true;
    if (window_oob(x1, y1, hw, hh, nc, nr) || window_oob(x2, y2, hw, hh, nc, nr)) { lvl_status = KLT_OOB; iterating = false; }

    float t_i[4], t_gx[4], t_gy[4];                            // previous frame, in ITS slot alignment m1
    float u_i[4], u_gx[4], u_gy[4];                            // the same, shifted to the new frame's alignment
    Row5 c_i, c_gx, c_gy;
    c_i.x = c_i.y = c_i.z = c_i.w = c_i.n = 0.f; c_gx = c_i; c_gy = c_i;
#pragma unroll
    for (int t = 0; t < 4; ++t) { t_i[t] = t_gx[t] = t_gy[t] = 0.f; u_i[t] = u_gx[t] = u_gy[t] = 0.f; }
    int c_xt = -1000000, c_yt = -1000000, m1 = 0, m2 = 0;
    bool vs[4] = {false, false, false, false};                // slot t holds a window sample
    if (iterating) {
      // all 6 loads of the level are issued before the first use
      const int xt1 = (int)x1, yt1 = (int)y1, xt2 = (int)x2, yt2 = (int)y2;
      m1 = (xt1 - 3) & 3; m2 = (xt2 - 3) & 3;
      const int o1 = (yt1 - 3 + r) * pitch + ((xt1 - 3) & ~3) + 4 * q;
      const int o2 = (yt2 - 3 + r) * pitch + ((xt2 - 3) & ~3) + 4 * q;
      const Row5 r_i = row5_load(p1.img[l], o1, ld, pol1), r_gx = row5_load(p1.gx[l], o1, ld, pol1),
                 r_gy = row5_load(p1.gy[l], o1, ld, pol1);
      c_i = row5_load(i2, o2, ld, pol2); c_gx = row5_load(gx2, o2, ld, pol2); c_gy = row5_load(gy2, o2, ld, pol2);
      c_xt = xt2; c_yt = yt2;
      const float ax = x1 - (float)xt1, ay = y1 - (float)yt1;
      row5_interp(r_i, ax, ay, t_i);
      row5_interp(r_gx, ax, ay, t_gx);
      row5_interp(r_gy, ax, ay, t_gy);
      row_shift(t_i, m2 - m1, lane, u_i); row_shift(t_gx, m2 - m1, lane, u_gx); row_shift(t_gy, m2 - m1, lane, u_gy);
#pragma unroll
      for (int t = 0; t < 4; ++t) { const int w = 4 * q + t - m2; vs[t] = r < 7 && ld && w >= 0 && w <= 6; }
    }

    float dx = 0.0f, dy = 0.0f;
    while (iterating) {                                       // warp uniform
      const int xt = (int)x2, yt = (int)y2;
      if (xt != c_xt || yt != c_yt) {                         // the integer footprint moved: re-read it
        const int o2 = (yt - 3 + r) * pitch + ((xt - 3) & ~3) + 4 * q;
        c_i = row5_load(i2, o2, ld, pol2); c_gx = row5_load(gx2, o2, ld, pol2); c_gy = row5_load(gy2, o2, ld, pol2);
        c_xt = xt; c_yt = yt;
        const int m = (xt - 3) & 3;
        if (m != m2) {
          m2 = m;
          row_shift(t_i, m2 - m1, lane, u_i); row_shift(t_gx, m2 - m1, lane, u_gx); row_shift(t_gy, m2 - m1, lane, u_gy);
#pragma unroll
          for (int t = 0; t < 4; ++t) { const int w = 4 * q + t - m2; vs[t] = r < 7 && ld && w >= 0 && w <= 6; }
        }
      }
      const float ax = x2 - (float)xt, ay = y2 - (float)yt;
      float s_i[4], s_gx[4], s_gy[4];
      row5_interp(c_i, ax, ay, s_i);
      row5_interp(c_gx, ax, ay, s_gx);
      row5_interp(c_gy, ax, ay, s_gy);
      float gxx = 0.0f, gxy = 0.0f, gyy = 0.0f, ex = 0.0f, ey = 0.0f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float df = vs[t] ? u_i[t] - s_i[t] : 0.0f, sx = vs[t] ? u_gx[t] + s_gx[t] : 0.0f,
                    sy = vs[t] ? u_gy[t] + s_gy[t] : 0.0f;
        gxx = fmaf(sx, sx, gxx); gxy = fmaf(sx, sy, gxy); gyy = fmaf(sy, sy, gyy);
        ex = fmaf(df, sx, ex); ey = fmaf(df, sy, ey);
      }
      gxx = warp_sum32(gxx); gxy = warp_sum32(gxy); gyy = warp_sum32(gyy);
      ex = warp_sum32(ex); ey = warp_sum32(ey);
      ex *= a.step_factor; ey *= a.step_factor;
      const float det = gxx * gyy - gxy * gxy;
      if (det < a.min_determinant) { lvl_status = KLT_SMALL_DET; break; }
      const float inv = __frcp_rn(det);
      dx = (gyy * ex - gxy * ey) * inv;
      dy = (gxx * ey - gxy * ex) * inv;
      x2 += dx; y2 += dy;
      ++iteration;
      const bool again = (fabsf(dx) >= a.min_displacement || fabsf(dy) >= a.min_displacement) &&
                         iteration < a.max_iterations;
      if (!again) break;
      if (window_oob(x2, y2, hw, hh, nc, nr)) { lvl_status = KLT_OOB; break; }
    }

    // after the loop (:459-474): bounds of the final position, then the residue
    if (lvl_status == KLT_TRACKED && window_oob(x2, y2, hw, hh, nc, nr)) lvl_status = KLT_OOB;
    if (lvl_status == KLT_TRACKED) {
      const int xt = (int)x2, yt = (int)y2;
      if (xt != c_xt || yt != c_yt) {
        c_i = row5_load(i2, (yt - 3 + r) * pitch + ((xt - 3) & ~3) + 4 * q, ld, pol2);
        const int m = (xt - 3) & 3;
        if (m != m2) {
          m2 = m;
          row_shift(t_i, m2 - m1, lane, u_i);
#pragma unroll
          for (int t = 0; t < 4; ++t) { const int w = 4 * q + t - m2; vs[t] = r < 7 && ld && w >= 0 && w <= 6; }
        }
      }
      float s_i[4];
      row5_interp(c_i, x2 - (float)xt, y2 - (float)yt, s_i);
      float sum = 0.0f;
#pragma unroll
      for (int t = 0; t < 4; ++t) sum += vs[t] ? fabsf(u_i[t] - s_i[t]) : 0.0f;
      sum = warp_sum32(sum);
      if (sum * (1.0f / 49.0f) > a.max_residue) lvl_status = KLT_LARGE_RESIDUE;
    }
    int v;                                                   // return value of _trackFeature (:479-484)
    if (lvl_status == KLT_SMALL_DET) v = KLT_SMALL_DET;
    else if (lvl_status == KLT_OOB) v = KLT_OOB;
    else if (lvl_status == KLT_LARGE_RESIDUE) v = KLT_LARGE_RESIDUE;
    else if (iteration >= a.max_iterations) v = KLT_MAX_ITERATIONS;
    else v = KLT_TRACKED;
    status = v;
    xout = x2; yout = y2;
    if (v == KLT_SMALL_DET || v == KLT_OOB) break;           // :1378
  }

  if (lane == 0) {                                           // record (:1383-1437)
    const bool outside = (xout < (float)a.borderx || xout > (float)(a.ncols - 1 - a.borderx) ||
                          yout < (float)a.bordery || yout > (float)(a.nrows - 1 - a.bordery));
    const size_t o = (size_t)f * io.ostride;
    if (status == KLT_OOB || outside) { io.ox[o] = -1.0f; io.oy[o] = -1.0f; io.oval[o] = KLT_OOB; }
    else if (status != KLT_TRACKED) { io.ox[o] = -1.0f; io.oy[o] = -1.0f; io.oval[o] = status; }
    else { io.ox[o] = xout; io.oy[o] = yout; io.oval[o] = KLT_TRACKED; }
  }
}
