// klt_fused.cuh -- TMA-staged, fused frame kernels (included by klt_dev.cu).
//
// l0_fused_kernel: one launch turns a u8 frame into level 0 of all three
// pyramids.  Replaces _KLTToFloatImage + _KLTComputeSmoothedImage +
// _KLTComputeGradients at level 0 (reference src/V1/convolve.c:37-53,273-314 as
// called from trackFeatures.c:1311-1321) without the float image, the
// horizontal-pass temporaries or the smoothed image ever being re-read from HBM:
// algorithmic traffic = 1 B read + 12 B written per pixel.
//
// Tile = 64x64 outputs, 256 threads, 4 stages in shared memory:
//   TMA   u8 box 96x74 at (x0-16, y0-5)         (zero fill outside the image)
//   A     horizontal Gaussian  -> Hs  [74][72]   (8 outputs / thread, one LDS.128 of 16 px)
//   B     vertical   Gaussian  -> L0  [70][72]   (4 cols x 5 rows / thread) + float4 stores of L0
//   C     horizontal DoG and G -> Hd,Hg [70][64] (8 outputs / thread)
//   D     vertical   G and DoG -> gx, gy         (4 cols x 8 rows / thread, float4 stores)
// Buffers have their column origin at x0-4 (16 B aligned) and row pitches chosen
// so that the 128-bit shared accesses of a quarter warp fall in distinct banks.
// The zero bands of the separable passes (convolve.c:164-178,216-237) are applied
// only by tiles that touch the image border (BORDER template flag).
#pragma once

#include <cuda.h>          // CUtensorMap (types only; the encoder is fetched at run time)

// ---- TMA / mbarrier primitives (sm_90+ PTX) ------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "KLT_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra KLT_DONE;\n\t"
      "bra KLT_WAIT;\n\t"
      "KLT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(phase)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int cx, int cy,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<unsigned long long>(map)), "r"(cx), "r"(cy), "r"(smem_u32(bar))
      : "memory");
}

// u8 -> f32 without the conversion pipe: 0x4B0000bb is the float 2^23 + bb.
__device__ __forceinline__ float u8_to_float(unsigned word, unsigned byte_sel) {
  return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7440u + byte_sel)) - 8388608.0f;
}

struct L0Geo {
  static constexpr int TX = 64, TY = 64;
  static constexpr int RS = 2, RG = 3;                 // smoothing / gradient radii
  static constexpr int U8_W = 96, U8_H = TY + 2 * (RS + RG);          // 96 x 74 bytes, cols <-> x0-16+c
  static constexpr int HS_P = 76, HS_H = U8_H;                         // cols <-> x0-4+c
  static constexpr int L0_P = 76, L0_H = TY + 2 * RG;                  // 70 rows <-> y0-3+r
  static constexpr int HG_P = 68, HG_H = L0_H;                         // cols <-> x0+c
  static constexpr int OFF_U8 = 0;
  static constexpr int OFF_HS = U8_W * U8_H;                           // 7104, 16 B aligned
  static constexpr int OFF_L0 = OFF_HS + HS_H * HS_P * 4;
  static constexpr int OFF_HD = OFF_HS;                                // Hd reuses Hs
  static constexpr int OFF_HG = OFF_L0 + L0_H * L0_P * 4;
  static constexpr int OFF_BAR = OFF_HG + HG_H * HG_P * 4;
  static constexpr int SMEM = OFF_BAR + 16;
};

// vertical 7-tap pass of 4 columns x 8 output rows from a shared [.][HG_P] buffer, scatter
// form: every loaded row feeds the (up to 7) outputs it belongs to, taps in increasing order.
template <bool EXACT, bool BORDER>
__device__ __forceinline__ void l0_stage_d(const float* __restrict__ src, const TapsR& tk,
                                           float* __restrict__ out, int opitch, int xg, int yg0,
                                           int W, int H) {
  constexpr int PY = 8, R = L0Geo::RG, P = L0Geo::HG_P;
  float4 acc[PY];
#pragma unroll
  for (int q = 0; q < PY; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < PY + 2 * R; ++i) {
    const float4 v = *reinterpret_cast<const float4*>(src + i * P);
#pragma unroll
    for (int q = 0; q < PY; ++q) {
      const int m = i - q;
      if (m >= 0 && m <= 2 * R) {
        acc[q].x = mac<EXACT>(acc[q].x, v.x, tk.k[m]);
        acc[q].y = mac<EXACT>(acc[q].y, v.y, tk.k[m]);
        acc[q].z = mac<EXACT>(acc[q].z, v.z, tk.k[m]);
        acc[q].w = mac<EXACT>(acc[q].w, v.w, tk.k[m]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < PY; ++q) {
    const int yg = yg0 + q;
    if (BORDER) {
      if (yg < R || yg >= H - R) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (xg < W && yg < H) *reinterpret_cast<float4*>(out + (size_t)yg * opitch + xg) = acc[q];
    } else {
      *reinterpret_cast<float4*>(out + (size_t)yg * opitch + xg) = acc[q];
    }
  }
}

template <bool EXACT, bool BORDER>
__device__ __forceinline__ void l0_fused_tile(unsigned char* smem, const CUtensorMap* map, int W, int H,
                                              const TapsR& ts, const TapsR& tg, const TapsR& td,
                                              float* __restrict__ out_img, float* __restrict__ out_gx,
                                              float* __restrict__ out_gy, int opitch, int x0, int y0) {
  using G = L0Geo;
  constexpr int RS = G::RS, RG = G::RG;
  unsigned char* sU8 = smem + G::OFF_U8;
  float* sHs = reinterpret_cast<float*>(smem + G::OFF_HS);
  float* sL0 = reinterpret_cast<float*>(smem + G::OFF_L0);
  float* sHd = reinterpret_cast<float*>(smem + G::OFF_HD);
  float* sHg = reinterpret_cast<float*>(smem + G::OFF_HG);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + G::OFF_BAR);
  const int tid = threadIdx.x;

  if (tid == 0) {
    mbar_init(bar, 1);
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar, G::U8_W * G::U8_H);
    // the innermost TMA start coordinate must be a multiple of 16 bytes (measured on B200:
    // x0-8 raises 'illegal instruction'), hence the 16-column left margin of the u8 tile
    tma_load_2d(sU8, map, x0 - 16, y0 - (RS + RG), bar);
  }
  mbar_wait(bar, 0);

  // lane -> (row, group) mapping shared by stages A and C: a quarter warp covers
  // 4 groups x 2 rows, which with the odd (in 16 B chunks) pitches is conflict free.
  const int lane = tid & 31, warp = tid >> 5;
  const int sub = lane & 3, rpar = (lane >> 2) & 1, half = (lane >> 3) & 1, rpair = lane >> 4;

  // ---- stage A: horizontal Gaussian, u8 -> Hs ------------------------------------------
  // 9 groups of 8 columns per row (72 cols <-> x0-4 .. x0+67); a warp takes 4 rows x 8 groups,
  // the 9th group of every row is handled by a tail loop.
  for (int rb = warp * 4; rb < G::HS_H; rb += 32) {
    const int r = rb + 2 * rpair + rpar;
    const int g = 4 * half + sub;
    if (r < G::HS_H) {
      const uint2 wa = *reinterpret_cast<const uint2*>(sU8 + r * G::U8_W + 8 * g + 8);
      const uint2 wb = *reinterpret_cast<const uint2*>(sU8 + r * G::U8_W + 8 * g + 16);
      const uint4 w = make_uint4(wa.x, wa.y, wb.x, wb.y);   // global cols x0-8+8g .. x0+7+8g
      float px[12];                                    // global cols x0-6+8g .. x0+5+8g
      px[0] = u8_to_float(w.x, 2); px[1] = u8_to_float(w.x, 3);
      px[2] = u8_to_float(w.y, 0); px[3] = u8_to_float(w.y, 1);
      px[4] = u8_to_float(w.y, 2); px[5] = u8_to_float(w.y, 3);
      px[6] = u8_to_float(w.z, 0); px[7] = u8_to_float(w.z, 1);
      px[8] = u8_to_float(w.z, 2); px[9] = u8_to_float(w.z, 3);
      px[10] = u8_to_float(w.w, 0); px[11] = u8_to_float(w.w, 1);
      float o[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float acc = 0.0f;
#pragma unroll
        for (int m = 0; m < 2 * RS + 1; ++m) acc = mac<EXACT>(acc, px[q + m], ts.k[m]);
        if (BORDER) {
          const int xg = x0 - 4 + 8 * g + q;
          if (xg < RS || xg >= W - RS) acc = 0.0f;
        }
        o[q] = acc;
      }
      float4* dst = reinterpret_cast<float4*>(sHs + r * G::HS_P + 8 * g);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
  // 9th group (cols 64..71): one item per row
  if (tid < G::HS_H) {
    const int r = tid, g = 8;
    const uint2 wa = *reinterpret_cast<const uint2*>(sU8 + r * G::U8_W + 8 * g + 8);
    const uint2 wb = *reinterpret_cast<const uint2*>(sU8 + r * G::U8_W + 8 * g + 16);
    const uint4 w = make_uint4(wa.x, wa.y, wb.x, wb.y);
    float px[12];
    px[0] = u8_to_float(w.x, 2); px[1] = u8_to_float(w.x, 3);
    px[2] = u8_to_float(w.y, 0); px[3] = u8_to_float(w.y, 1);
    px[4] = u8_to_float(w.y, 2); px[5] = u8_to_float(w.y, 3);
    px[6] = u8_to_float(w.z, 0); px[7] = u8_to_float(w.z, 1);
    px[8] = u8_to_float(w.z, 2); px[9] = u8_to_float(w.z, 3);
    px[10] = u8_to_float(w.w, 0); px[11] = u8_to_float(w.w, 1);
    float o[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float acc = 0.0f;
#pragma unroll
      for (int m = 0; m < 2 * RS + 1; ++m) acc = mac<EXACT>(acc, px[q + m], ts.k[m]);
      if (BORDER) {
        const int xg = x0 - 4 + 8 * g + q;
        if (xg < RS || xg >= W - RS) acc = 0.0f;
      }
      o[q] = acc;
    }
    float4* dst = reinterpret_cast<float4*>(sHs + r * G::HS_P + 8 * g);
    dst[0] = make_float4(o[0], o[1], o[2], o[3]);
    dst[1] = make_float4(o[4], o[5], o[6], o[7]);
  }
  __syncthreads();

  // ---- stage B: vertical Gaussian, Hs -> L0 (shared + global) ------------------------------
  // 18 column groups of 4 (cols <-> x0-4+4j) x 14 row blocks of 5 (L0 rows <-> y0-3+rr)
  if (tid < 18 * 14) {
    const int blk = tid / 18, j = tid - blk * 18;
    constexpr int PY = 5;
    float4 acc[PY];
#pragma unroll
    for (int q = 0; q < PY; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* src = sHs + (blk * PY) * G::HS_P + 4 * j;      // Hs row = L0 row + (RS) - RS ...
#pragma unroll
    for (int i = 0; i < PY + 2 * RS; ++i) {
      // L0 row rr (tile) <-> y0-3+rr needs Hs rows (y0-5+..): Hs row index = rr + m, m = 0..2RS
      const float4 v = *reinterpret_cast<const float4*>(src + i * G::HS_P);
#pragma unroll
      for (int q = 0; q < PY; ++q) {
        const int m = i - q;
        if (m >= 0 && m <= 2 * RS) {
          acc[q].x = mac<EXACT>(acc[q].x, v.x, ts.k[m]);
          acc[q].y = mac<EXACT>(acc[q].y, v.y, ts.k[m]);
          acc[q].z = mac<EXACT>(acc[q].z, v.z, ts.k[m]);
          acc[q].w = mac<EXACT>(acc[q].w, v.w, ts.k[m]);
        }
      }
    }
    const int xg = x0 - 4 + 4 * j;
#pragma unroll
    for (int q = 0; q < PY; ++q) {
      const int rr = blk * PY + q;
      const int yg = y0 - RG + rr;
      if (BORDER) {
        if (yg < RS || yg >= H - RS) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      *reinterpret_cast<float4*>(sL0 + rr * G::L0_P + 4 * j) = acc[q];
      if (j >= 1 && j <= 16 && rr >= RG && rr < RG + G::TY) {
        if (!BORDER || (xg < W && yg < H))
          *reinterpret_cast<float4*>(out_img + (size_t)yg * opitch + xg) = acc[q];
      }
    }
  }
  __syncthreads();

  // ---- stage C: horizontal DoG and Gaussian, L0 -> Hd, Hg ----------------------------------
  // 8 groups of 8 output columns (x0+8g ..) per L0 row; window = L0 tile cols 8g .. 8g+15
  for (int rb = warp * 4; rb < G::L0_H; rb += 32) {
    const int r = rb + 2 * rpair + rpar;
    const int g = 4 * half + sub;
    if (r < G::L0_H) {
      float win[16];
      const float4* p = reinterpret_cast<const float4*>(sL0 + r * G::L0_P + 8 * g);
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const float4 t = p[v];
        win[4 * v] = t.x; win[4 * v + 1] = t.y; win[4 * v + 2] = t.z; win[4 * v + 3] = t.w;
      }
      float od[8], og[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float a = 0.0f, b = 0.0f;
#pragma unroll
        for (int m = 0; m < 2 * RG + 1; ++m) a = mac<EXACT>(a, win[q + 1 + m], td.k[m]);
#pragma unroll
        for (int m = 0; m < 2 * RG + 1; ++m) b = mac<EXACT>(b, win[q + 1 + m], tg.k[m]);
        if (BORDER) {
          const int xg = x0 + 8 * g + q;
          if (xg < RG || xg >= W - RG) { a = 0.0f; b = 0.0f; }
        }
        od[q] = a; og[q] = b;
      }
      float4* dd = reinterpret_cast<float4*>(sHd + r * G::HG_P + 8 * g);
      float4* dg = reinterpret_cast<float4*>(sHg + r * G::HG_P + 8 * g);
      dd[0] = make_float4(od[0], od[1], od[2], od[3]);
      dd[1] = make_float4(od[4], od[5], od[6], od[7]);
      dg[0] = make_float4(og[0], og[1], og[2], og[3]);
      dg[1] = make_float4(og[4], og[5], og[6], og[7]);
    }
  }
  __syncthreads();

  // ---- stage D: vertical passes, Hd -> gx (Gaussian), Hg -> gy (DoG) -------------------------
  // threads 0..127 produce gx, 128..255 gy (warp uniform); each 4 columns x 8 rows.
  {
    const int t = tid & 127;
    const int blk = t >> 4, j = t & 15;             // 8 row blocks x 16 column groups
    if (tid < 128)
      l0_stage_d<EXACT, BORDER>(sHd + (blk * 8) * G::HG_P + 4 * j, tg, out_gx, opitch, x0 + 4 * j,
                                y0 + blk * 8, W, H);
    else
      l0_stage_d<EXACT, BORDER>(sHg + (blk * 8) * G::HG_P + 4 * j, td, out_gy, opitch, x0 + 4 * j,
                                y0 + blk * 8, W, H);
  }
}

template <bool EXACT>
__global__ void __launch_bounds__(256, 3)
l0_fused_kernel(const __grid_constant__ CUtensorMap map, int W, int H, TapsR ts, TapsR tg, TapsR td,
                float* __restrict__ out_img, float* __restrict__ out_gx, float* __restrict__ out_gy,
                int opitch) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int x0 = blockIdx.x * L0Geo::TX, y0 = blockIdx.y * L0Geo::TY;
  // tiles whose 8-pixel margin stays inside the image never meet a zero band
  const bool border = (x0 < 8) || (y0 < 8) || (x0 + L0Geo::TX + 8 > W) || (y0 + L0Geo::TY + 8 > H);
  if (border)
    l0_fused_tile<EXACT, true>(smem_raw, &map, W, H, ts, tg, td, out_img, out_gx, out_gy, opitch, x0, y0);
  else
    l0_fused_tile<EXACT, false>(smem_raw, &map, W, H, ts, tg, td, out_img, out_gx, out_gy, opitch, x0, y0);
}
