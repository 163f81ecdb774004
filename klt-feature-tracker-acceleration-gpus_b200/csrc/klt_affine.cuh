// klt_affine.cuh -- the affine consistency check of KLTTrackFeatures on the device
// (reference src/V1/trackFeatures.c:506-1224 "_am_*" helpers, :1438-1497 the per-feature logic).
//
// After the translation tracker has placed a feature in the new frame, the reference compares the
// window around it with a template saved when the feature was first tracked: a second Newton
// iteration over a 15x15 window whose unknowns are a translation (check 0), a similarity (check 1:
// 4 unknowns) or a full affine map (check 2: 6 unknowns) plus the translation; the feature is
// dropped when that iteration leaves the image, drifts more than affine_max_displacement_differ or
// ends with a residue above affine_max_residue.  The refined position is NOT written back
// (:1490-1491); only the status and the map A persist.
//
// One warp per feature.  The arithmetic is the reference's, operation for operation and always in
// its order (separately rounded multiply / add: __fmul_rn / __fadd_rn keep the compiler from
// contracting them): the window samples go to shared memory, then every normal-equation entry is a
// sequential raster-order sum owned by one lane (27 lanes for the 6x6 system + right-hand side),
// lane 0 runs the Gauss-Jordan elimination with full pivoting in shared memory.  The result is
// bit-identical to the CPU reference for identical pyramids and positions (tests/test_gpu_affine.py).
#pragma once

#define AMUL(a, b) __fmul_rn((a), (b))
#define AADD(a, b) __fadd_rn((a), (b))
#define ASUB(a, b) __fsub_rn((a), (b))

struct AffState {              // the aff_* members of KLT_FeatureRec, per feature (32 B)
  int   has;                   // a template is held (aff_img != NULL)
  float aff_x, aff_y, Axx, Ayx, Axy, Ayy;
  int   flags;                 // out: 1 = template created in this call, 2 = template released
};

struct AffArgs {
  int   check, aw, ah, max_iterations;
  float max_residue, th_aff, mdd;
  float step_factor, small, th;
  int   nlevels; float ss;
  int   ncols, nrows, pitch;                 // level 0
  int   resident;                            // the state array lives on the device across frames (sequence call)
  int   lighting;                            // tc->lighting_insensitive: check 0 refines on gain / bias normalised windows
  const float *i1, *gx1, *gy1;               // previous frame, level 0: template source
  const float *i2, *gx2, *gy2;               // new frame, level 0
};

// normal-equation bookkeeping of the 6x6 system: lane -> (row, col, coefficient, gradient product)
// coefficient: 0 xx, 1 xy, 2 yy, 3 x, 4 y, 5 one;  product: 0 gxx, 1 gxy, 2 gyy, 3 diff*gx, 4 diff*gy
__constant__ unsigned char kAff6[27][4] = {
  {0, 0, 0, 0}, {0, 1, 0, 1}, {0, 2, 1, 0}, {0, 3, 1, 1}, {0, 4, 3, 0}, {0, 5, 3, 1},
  {1, 1, 0, 2}, {1, 2, 1, 1}, {1, 3, 1, 2}, {1, 4, 3, 1}, {1, 5, 3, 2},
  {2, 2, 2, 0}, {2, 3, 2, 1}, {2, 4, 4, 0}, {2, 5, 4, 1},
  {3, 3, 2, 2}, {3, 4, 4, 1}, {3, 5, 4, 2},
  {4, 4, 5, 0}, {4, 5, 5, 1}, {5, 5, 5, 2},
  // right-hand side (:806-838): e0 dgx*i, e1 dgy*i, e2 dgx*j, e3 dgy*j, e4 dgx, e5 dgy
  {6, 0, 3, 3}, {6, 1, 3, 4}, {6, 2, 4, 3}, {6, 3, 4, 4}, {6, 4, 5, 3}, {6, 5, 5, 4}};
// 4x4 system (:846-928): operands 0 u = x*gx + y*gy, 1 w = x*gy - y*gx, 2 gx, 3 gy
__constant__ unsigned char kAff4[10][4] = {
  {0, 0, 0, 0}, {0, 1, 0, 1}, {0, 2, 0, 2}, {0, 3, 0, 3}, {1, 1, 1, 1}, {1, 2, 1, 2}, {1, 3, 1, 3},
  {2, 2, 2, 2}, {2, 3, 2, 3}, {3, 3, 3, 3}};

__device__ __forceinline__ float aff_bilinear(const float* __restrict__ img, int pitch, float x, float y) {
  const Bilin b = bilin_setup<true>(x, y, pitch);
  return bilin_fetch<true>(img, pitch, b);
}

// trackFeatures.c:546-604, n x n, one right-hand side, in shared memory; run by one lane
__device__ int aff_gauss_jordan(float* a /* [6][6] */, int n, float* b) {
  int ipiv[6];
  int col = 0, row = 0;
  for (int j = 0; j < n; ++j) ipiv[j] = 0;
  for (int i = 0; i < n; ++i) {
    float big = 0.0f;
    for (int j = 0; j < n; ++j)
      if (ipiv[j] != 1)
        for (int k = 0; k < n; ++k) {
          if (ipiv[k] == 0) {
            if (fabsf(a[j * 6 + k]) >= big) { big = fabsf(a[j * 6 + k]); row = j; col = k; }
          } else if (ipiv[k] > 1) return KLT_SMALL_DET;
        }
    ++ipiv[col];
    if (row != col) {
      for (int l = 0; l < n; ++l) { const float t = a[row * 6 + l]; a[row * 6 + l] = a[col * 6 + l]; a[col * 6 + l] = t; }
      const float t = b[row]; b[row] = b[col]; b[col] = t;
    }
    if (a[col * 6 + col] == 0.0f) return KLT_SMALL_DET;
    const float pivinv = __fdiv_rn(1.0f, a[col * 6 + col]);
    a[col * 6 + col] = 1.0f;
    for (int l = 0; l < n; ++l) a[col * 6 + l] = AMUL(a[col * 6 + l], pivinv);
    b[col] = AMUL(b[col], pivinv);
    for (int ll = 0; ll < n; ++ll)
      if (ll != col) {
        const float dum = a[ll * 6 + col];
        a[ll * 6 + col] = 0.0f;
        for (int l = 0; l < n; ++l) a[ll * 6 + l] = ASUB(a[ll * 6 + l], AMUL(a[col * 6 + l], dum));
        b[ll] = ASUB(b[ll], AMUL(b[col], dum));
      }
  }
  return KLT_TRACKED;
}

__global__ void __launch_bounds__(128)
affine_check_kernel(AffArgs a, int n, const float* __restrict__ x0, const float* __restrict__ y0,
                    const int* __restrict__ val0, float* x, float* y, int* val,
                    AffState* st, float* tmpl) {
  extern __shared__ float s_aff[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int f = blockIdx.x * (blockDim.x >> 5) + wib;
  if (f >= n) return;
  if (val0[f] < 0) return;                                    // not tracked in this call (:1346)
  const int npix = a.aw * a.ah, tw = a.aw + 2, th = a.ah + 2, tsz = tw * th;
  float* sd = s_aff + (size_t)wib * (3 * npix + 48);
  float* sx = sd + npix;
  float* sy = sx + npix;
  float* sT = sy + npix;                                      // 36
  float* sa = sT + 36;                                        // 6 (+ padding)
  AffState s = st[f];
  if (a.resident && val0[f] > 0) {
    // a slot KLTReplaceLostFeatures refilled since the last frame: a new feature, its affine members
    // start over (selectGoodFeatures.c:514-541) -- in the per-call API the host list carries the reset
    s.has = 0; s.aff_x = -1.0f; s.aff_y = -1.0f; s.Axx = 1.0f; s.Ayx = 0.0f; s.Axy = 0.0f; s.Ayy = 1.0f;
  }
  float* t_img = tmpl + (size_t)f * 3 * tsz;
  float* t_gx = t_img + tsz;
  float* t_gy = t_gx + tsz;
  __syncwarp();                                               // every lane has read st[f]

  if (val[f] != KLT_TRACKED) {       // lost by the translation tracker: the template goes (:1383-1431)
    if (lane == 0 && (s.has || (a.resident && val0[f] > 0))) { AffState o = s; o.has = 0; o.flags = 2; st[f] = o; }
    return;
  }

  if (!s.has) {
    // :1446-1457 first successful track: keep the (aw+2) x (ah+2) neighbourhood of the OLD position
    // in the previous frame's level 0 (image and gradients) as the template
    float xloc = x0[f], yloc = y0[f];
    for (int l = 0; l < a.nlevels; ++l) { xloc = __fdiv_rn(xloc, a.ss); yloc = __fdiv_rn(yloc, a.ss); }
    for (int l = 0; l < a.nlevels; ++l) { xloc = AMUL(xloc, a.ss); yloc = AMUL(yloc, a.ss); }
    const int xi = (int)xloc, yi = (int)yloc, hwt = tw / 2, hht = th / 2;
    for (int k = lane; k < tsz; k += 32) {
      int px = xi + (k % tw) - hwt, py = yi + (k / tw) - hht;
      px = px < 0 ? 0 : (px > a.ncols - 1 ? a.ncols - 1 : px);        // the reference asserts the range
      py = py < 0 ? 0 : (py > a.nrows - 1 ? a.nrows - 1 : py);
      const size_t off = (size_t)py * a.pitch + px;
      t_img[k] = a.i1[off]; t_gx[k] = a.gx1[off]; t_gy[k] = a.gy1[off];
    }
    if (lane == 0) {
      AffState o = s;
      o.has = 1; o.flags = 1;
      o.aff_x = AADD(ASUB(xloc, (float)xi), (float)((a.aw + 2) / 2));
      o.aff_y = AADD(ASUB(yloc, (float)yi), (float)((a.ah + 2) / 2));
      st[f] = o;
    }
    return;
  }

  // :1458-1493 / :952-1224 refine against the template
  const int hw = a.aw / 2, hh = a.ah / 2;
  const float fhw = (float)hw, fhh = (float)hh, nhw = (float)(-hw), nhh = (float)(-hh);
  const int nc1 = tw, nr1 = th, nc2 = a.ncols, nr2 = a.nrows;
  const float x1 = s.aff_x, y1 = s.aff_y;
  float x2 = x[f], y2 = y[f];
  const float old_x2 = x2, old_y2 = y2;
  float Axx = s.Axx, Ayx = s.Ayx, Axy = s.Axy, Ayy = s.Ayy;
  float dx = 0.0f, dy = 0.0f;
  int iteration = 0, status = KLT_TRACKED;
  bool convergence = false;
  const float eps1 = 1.001f;

  // this lane's normal-equation entry
  int e_row = 7, e_col = 0, e_a = 0, e_b = 0;
  if (a.check == 2 && lane < 27) { e_row = kAff6[lane][0]; e_col = kAff6[lane][1]; e_a = kAff6[lane][2]; e_b = kAff6[lane][3]; }
  if (a.check == 1 && lane < 10) { e_row = kAff4[lane][0]; e_col = kAff4[lane][1]; e_a = kAff4[lane][2]; e_b = kAff4[lane][3]; }
  if (a.check == 1 && lane >= 10 && lane < 14) { e_row = 6; e_col = lane - 10; }

  do {
    float ul_x = 0.f, ul_y = 0.f, ll_x = 0.f, ll_y = 0.f, ur_x = 0.f, ur_y = 0.f, lr_x = 0.f, lr_y = 0.f;
    if (a.check == 0) {
      if (window_oob(x1, y1, hw, hh, nc1, nr1) || window_oob(x2, y2, hw, hh, nc2, nr2)) { status = KLT_OOB; break; }
    } else {
      ul_x = AADD(AADD(AMUL(Axx, nhw), AMUL(Axy, fhh)), x2); ul_y = AADD(AADD(AMUL(Ayx, nhw), AMUL(Ayy, fhh)), y2);
      ll_x = AADD(AADD(AMUL(Axx, nhw), AMUL(Axy, nhh)), x2); ll_y = AADD(AADD(AMUL(Ayx, nhw), AMUL(Ayy, nhh)), y2);
      ur_x = AADD(AADD(AMUL(Axx, fhw), AMUL(Axy, fhh)), x2); ur_y = AADD(AADD(AMUL(Ayx, fhw), AMUL(Ayy, fhh)), y2);
      lr_x = AADD(AADD(AMUL(Axx, fhw), AMUL(Axy, nhh)), x2); lr_y = AADD(AADD(AMUL(Ayx, fhw), AMUL(Ayy, nhh)), y2);
      const float fc = (float)nc2, fr = (float)nr2;
      if (window_oob(x1, y1, hw, hh, nc1, nr1) ||
          ul_x < 0.0f || ASUB(fc, ul_x) < eps1 || ll_x < 0.0f || ASUB(fc, ll_x) < eps1 ||
          ur_x < 0.0f || ASUB(fc, ur_x) < eps1 || lr_x < 0.0f || ASUB(fc, lr_x) < eps1 ||
          ul_y < 0.0f || ASUB(fr, ul_y) < eps1 || ll_y < 0.0f || ASUB(fr, ll_y) < eps1 ||
          ur_y < 0.0f || ASUB(fr, ur_y) < eps1 || lr_y < 0.0f || ASUB(fr, lr_y) < eps1) { status = KLT_OOB; break; }
    }
    if (a.check == 0 && a.lighting) {
      // :1024-1028 -> :125-220: gain / bias normalised windows, template against frame.  Pass 1 leaves
      // the raw samples g1 (template) in sd and g2 (frame) in sx; four lanes sum them sequentially in
      // raster order (sum g1, sum g2, sum g1^2, sum g2^2: the reference's order and rounding); pass 2
      // forms  g1 - g2 * alpha - belta  and  grad1 + grad2 * alpha_g  (whose gain is the reference's
      // sqrt(mean g1 / mean g2), :202).
      for (int k = lane; k < npix; k += 32) {
        const float fi = (float)(k % a.aw - hw), fj = (float)(k / a.aw - hh);
        sd[k] = aff_bilinear(t_img, tw, AADD(x1, fi), AADD(y1, fj));
        sx[k] = aff_bilinear(a.i2, a.pitch, AADD(x2, fi), AADD(y2, fj));
      }
      __syncwarp();
      float acc = 0.0f;
      if (lane < 4) {
        const float* A = (lane & 1) ? sx : sd;
        if (lane < 2) { for (int k = 0; k < npix; ++k) acc = AADD(acc, A[k]); }
        else          { for (int k = 0; k < npix; ++k) acc = AADD(acc, AMUL(A[k], A[k])); }
      }
      const float s1 = __shfl_sync(0xffffffffu, acc, 0), s2 = __shfl_sync(0xffffffffu, acc, 1);
      const float q1 = __shfl_sync(0xffffffffu, acc, 2), q2 = __shfl_sync(0xffffffffu, acc, 3);
      const float fn = (float)npix;
      const float alpha = (float)sqrt((double)__fdiv_rn(__fdiv_rn(q1, fn), __fdiv_rn(q2, fn)));
      const float m1 = __fdiv_rn(s1, fn), m2 = __fdiv_rn(s2, fn);
      const float belta = ASUB(m1, AMUL(alpha, m2));
      const float alphag = (float)sqrt((double)__fdiv_rn(m1, m2));
      __syncwarp();
      for (int k = lane; k < npix; k += 32) {
        const float fi = (float)(k % a.aw - hw), fj = (float)(k / a.aw - hh);
        const float tx = AADD(x1, fi), ty = AADD(y1, fj), qx = AADD(x2, fi), qy = AADD(y2, fj);
        const float g1 = sd[k], g2 = sx[k];
        sd[k] = ASUB(ASUB(g1, AMUL(g2, alpha)), belta);
        sx[k] = AADD(aff_bilinear(t_gx, tw, tx, ty), AMUL(aff_bilinear(a.gx2, a.pitch, qx, qy), alphag));
        sy[k] = AADD(aff_bilinear(t_gy, tw, tx, ty), AMUL(aff_bilinear(a.gy2, a.pitch, qx, qy), alphag));
      }
    } else
    // window samples (:68-123 for check 0; :700-722 and :610-632 otherwise)
    for (int k = lane; k < npix; k += 32) {
      const float fi = (float)(k % a.aw - hw), fj = (float)(k / a.aw - hh);
      const float tx = AADD(x1, fi), ty = AADD(y1, fj);
      const float g1 = aff_bilinear(t_img, tw, tx, ty);
      if (a.check == 0) {
        const float qx = AADD(x2, fi), qy = AADD(y2, fj);
        sd[k] = ASUB(g1, aff_bilinear(a.i2, a.pitch, qx, qy));
        sx[k] = AADD(aff_bilinear(t_gx, tw, tx, ty), aff_bilinear(a.gx2, a.pitch, qx, qy));
        sy[k] = AADD(aff_bilinear(t_gy, tw, tx, ty), aff_bilinear(a.gy2, a.pitch, qx, qy));
      } else {
        const float mi = AADD(AMUL(Axx, fi), AMUL(Axy, fj)), mj = AADD(AMUL(Ayx, fi), AMUL(Ayy, fj));
        const float qx = AADD(x2, mi), qy = AADD(y2, mj);
        sd[k] = ASUB(g1, aff_bilinear(a.i2, a.pitch, qx, qy));
        sx[k] = aff_bilinear(a.gx2, a.pitch, qx, qy);
        sy[k] = aff_bilinear(a.gy2, a.pitch, qx, qy);
      }
    }
    __syncwarp();

    if (a.check == 0) {
      // :227-307 the 2x2 system of the translation tracker, on the template window
      float acc = 0.0f;
      if (lane < 5)
        for (int k = 0; k < npix; ++k) {
          const float wx = sx[k], wy = sy[k], df = sd[k];
          const float A = lane < 3 ? (lane == 2 ? wy : wx) : df;
          const float B = (lane == 0 || lane == 3) ? wx : wy;
          acc = AADD(acc, AMUL(A, B));
        }
      const float gxx = __shfl_sync(0xffffffffu, acc, 0), gxy = __shfl_sync(0xffffffffu, acc, 1);
      const float gyy = __shfl_sync(0xffffffffu, acc, 2);
      const float ex = AMUL(__shfl_sync(0xffffffffu, acc, 3), a.step_factor);
      const float ey = AMUL(__shfl_sync(0xffffffffu, acc, 4), a.step_factor);
      const float det = ASUB(AMUL(gxx, gyy), AMUL(gxy, gxy));
      if (det < a.small) status = KLT_SMALL_DET;
      else {
        dx = __fdiv_rn(ASUB(AMUL(gyy, ex), AMUL(gxy, ey)), det);
        dy = __fdiv_rn(ASUB(AMUL(gxx, ey), AMUL(gxy, ex)), det);
        status = KLT_TRACKED;
      }
      convergence = fabsf(dx) < a.th && fabsf(dy) < a.th;
      x2 = AADD(x2, dx); y2 = AADD(y2, dy);
    } else {
      float acc = 0.0f;
      if (e_row < 7) {
        for (int k = 0; k < npix; ++k) {
          const float gx = sx[k], gy = sy[k], df = sd[k];
          const float fx = (float)(k % a.aw - hw), fy = (float)(k / a.aw - hh);
          float term;
          if (a.check == 2) {
            const float xx = AMUL(fx, fx), xy = AMUL(fx, fy), yy = AMUL(fy, fy);
            const float A = e_a == 0 ? xx : e_a == 1 ? xy : e_a == 2 ? yy : e_a == 3 ? fx : e_a == 4 ? fy : 1.0f;
            const float B = e_b == 0 ? AMUL(gx, gx) : e_b == 1 ? AMUL(gx, gy) : e_b == 2 ? AMUL(gy, gy)
                          : e_b == 3 ? AMUL(df, gx) : AMUL(df, gy);
            term = AMUL(A, B);
          } else if (e_row < 6) {
            const float u = AADD(AMUL(fx, gx), AMUL(fy, gy)), w = ASUB(AMUL(fx, gy), AMUL(fy, gx));
            const float A = e_a == 0 ? u : e_a == 1 ? w : e_a == 2 ? gx : gy;
            const float B = e_b == 0 ? u : e_b == 1 ? w : e_b == 2 ? gx : gy;
            term = AMUL(A, B);
          } else {
            const float dgx = AMUL(df, gx), dgy = AMUL(df, gy);
            term = e_col == 0 ? AADD(AMUL(dgx, fx), AMUL(dgy, fy))
                 : e_col == 1 ? ASUB(AMUL(dgy, fx), AMUL(dgx, fy)) : e_col == 2 ? dgx : dgy;
          }
          acc = AADD(acc, term);
        }
        if (e_row == 6) sa[e_col] = AMUL(acc, 0.5f);
        else { sT[e_row * 6 + e_col] = acc; sT[e_col * 6 + e_row] = acc; }
      }
      __syncwarp();
      int st_solve = KLT_TRACKED;
      if (lane == 0) st_solve = aff_gauss_jordan(sT, a.check == 2 ? 6 : 4, sa);
      st_solve = __shfl_sync(0xffffffffu, st_solve, 0);
      __syncwarp();
      status = st_solve;
      if (a.check == 1) {
        Axx = AADD(Axx, sa[0]); Ayx = AADD(Ayx, sa[1]); Ayy = Axx; Axy = -Ayx;
        dx = sa[2]; dy = sa[3];
      } else {
        Axx = AADD(Axx, sa[0]); Ayx = AADD(Ayx, sa[1]); Axy = AADD(Axy, sa[2]); Ayy = AADD(Ayy, sa[3]);
        dx = sa[4]; dy = sa[5];
      }
      x2 = AADD(x2, dx); y2 = AADD(y2, dy);
      // :1162-1173 how far the window corners moved in this iteration
      ul_x = ASUB(ul_x, AADD(AADD(AMUL(Axx, nhw), AMUL(Axy, fhh)), x2)); ul_y = ASUB(ul_y, AADD(AADD(AMUL(Ayx, nhw), AMUL(Ayy, fhh)), y2));
      ll_x = ASUB(ll_x, AADD(AADD(AMUL(Axx, nhw), AMUL(Axy, nhh)), x2)); ll_y = ASUB(ll_y, AADD(AADD(AMUL(Ayx, nhw), AMUL(Ayy, nhh)), y2));
      ur_x = ASUB(ur_x, AADD(AADD(AMUL(Axx, fhw), AMUL(Axy, fhh)), x2)); ur_y = ASUB(ur_y, AADD(AADD(AMUL(Ayx, fhw), AMUL(Ayy, fhh)), y2));
      lr_x = ASUB(lr_x, AADD(AADD(AMUL(Axx, fhw), AMUL(Axy, nhh)), x2)); lr_y = ASUB(lr_y, AADD(AADD(AMUL(Ayx, fhw), AMUL(Ayy, nhh)), y2));
      convergence = fabsf(dx) < a.th && fabsf(dy) < a.th &&
                    fabsf(ul_x) < a.th_aff && fabsf(ul_y) < a.th_aff && fabsf(ll_x) < a.th_aff && fabsf(ll_y) < a.th_aff &&
                    fabsf(ur_x) < a.th_aff && fabsf(ur_y) < a.th_aff && fabsf(lr_x) < a.th_aff && fabsf(lr_y) < a.th_aff;
    }
    if (status == KLT_SMALL_DET) break;
    ++iteration;
  } while (!convergence && iteration < a.max_iterations);
  __syncwarp();

  // :1193-1200
  if (window_oob(x2, y2, hw, hh, nc2, nr2)) status = KLT_OOB;
  if (ASUB(x2, old_x2) > a.mdd || ASUB(y2, old_y2) > a.mdd) status = KLT_OOB;
  // :1203-1214 residue of the final alignment
  if (status == KLT_TRACKED) {
    for (int k = lane; k < npix; k += 32) {
      const float fi = (float)(k % a.aw - hw), fj = (float)(k / a.aw - hh);
      const float g1 = aff_bilinear(t_img, tw, AADD(x1, fi), AADD(y1, fj));
      float qx, qy;
      if (a.check == 0) { qx = AADD(x2, fi); qy = AADD(y2, fj); }
      else {
        qx = AADD(x2, AADD(AMUL(Axx, fi), AMUL(Axy, fj)));
        qy = AADD(y2, AADD(AMUL(Ayx, fi), AMUL(Ayy, fj)));
      }
      sd[k] = ASUB(g1, aff_bilinear(a.i2, a.pitch, qx, qy));
    }
    __syncwarp();
    float sum = 0.0f;
    if (lane == 0)
      for (int k = 0; k < npix; ++k) sum = AADD(sum, fabsf(sd[k]));
    sum = __shfl_sync(0xffffffffu, sum, 0);
    if (__fdiv_rn(sum, (float)npix) > a.max_residue) status = KLT_LARGE_RESIDUE;
  }

  if (lane == 0) {
    AffState o = s;
    o.Axx = Axx; o.Ayx = Ayx; o.Axy = Axy; o.Ayy = Ayy; o.flags = 0;
    val[f] = status;
    if (status != KLT_TRACKED) {                               // :1480-1489
      x[f] = -1.0f; y[f] = -1.0f;
      o.aff_x = -1.0f; o.aff_y = -1.0f; o.has = 0; o.flags = 2;
    }
    st[f] = o;
  }
}
