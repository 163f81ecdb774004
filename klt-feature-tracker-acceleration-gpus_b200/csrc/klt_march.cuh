// klt_march.cuh -- level 0 as column strips marched down by warp teams (included by klt_dev.cu).
//
// l0_march_kernel: u8 frame -> L0, gx0, gy0 (1 B read + 12 B written per pixel), the same arithmetic
// as l0_fused_kernel (klt_fused.cuh) in a different schedule.  Replaces _KLTToFloatImage +
// _KLTComputeSmoothedImage + _KLTComputeGradients of level 0 (reference src/V1/convolve.c:37-53,
// 137-242, 273-314 as sequenced by trackFeatures.c:1311-1321).
//
// The tile kernel pays four CTA-wide barriers per tile and recomputes a 10-row halo per 48-row tile.
// Here the unit of work is a column strip of 120 pixels x a segment of rows, walked top to bottom by a
// TEAM of three warps that never meets a CTA barrier:
//   producer  lane = 4 columns (strip + 4 halo columns each side).  Per frame row: one 64-bit load (its
//             own 4 pixels + one neighbour word, issued 5 rows ahead), the other neighbour's pixels by
//             one butterfly shuffle of the raw word, horizontal 5-tap Gaussian in registers, then the row
//             is SCATTERED into the five running vertical sums it belongs to (taps in increasing row
//             order == the reference's summation order).  The sum that completes is a row of L0: stored
//             to HBM and into a 28-row ring in shared memory.
//   gx, gy    two consumer warps, lane = 4 columns.  Per L0 row of the ring: three 128-bit shared loads
//             (12 columns), horizontal derivative (gx) or Gaussian (gy) 7-tap in registers, scatter into
//             seven running vertical sums, the completed row goes to HBM.
// Vertical passes therefore need no halo recomputation inside a segment (only 10 / 6 warm-up rows at
// its top), no intermediate but the L0 ring touches shared memory, and the only synchronisation is a
// pair of mbarriers per 7-row chunk of the ring between the three warps of a team (full: producer ->
// consumers, empty: consumers -> producer).  Teams are independent: 4 per CTA, 2 CTAs per SM.
//
// Border semantics as everywhere (klt_dev.cu header): outputs in the zero bands are written as zeros
// by the tasks that touch the image border; what those tasks read outside the image is clamped to the
// nearest pixel / row and only ever feeds zeroed outputs.  A zero band of a horizontal pass is applied
// to the completed row of the vertical pass that follows it (a column of zeros sums to zero), so each
// warp has ONE fix-up per row, behind a warp-uniform branch that interior tasks never take.
//
// What the first versions taught (ncu, 4K frame): (1) separate interior / border instantiations of
// the three roles made 88 KB of code: 45 % of the warp samples were "no instruction", 63 us.  (2) A
// load issued into a register-ring slot whose old value is still live lands in a temporary and is
// MOVed into the slot at the end of the row -- the move waits for the load: the DRAM latency of every
// row exposed, 52 us; the load is now issued after the last use of the slot's old value.  (3) After
// that the kernel is issue bound (70 % issue utilisation, a third of it FMAs): everything below is
// arranged to keep the per-row instruction count down.
#pragma once

struct MarchGeo {
  static constexpr int SWI = 120;               // interior columns of a strip (30 lanes x 4)
  static constexpr int LW = 128;                // L0 columns in the ring: xs-4 .. xs+123
  static constexpr int RS = 2, RG = FUSED_RG;   // smoothing / gradient radii
  static constexpr int PU = 2 * RS + 1;         // producer unroll = its running sums
  static constexpr int CH = 2 * RG + 1;         // ring chunk = 7 rows == the consumers' unroll
  static constexpr int NSLOT = 4;               // chunks in the ring
  static constexpr int TEAMS = 4;               // teams per CTA
  static constexpr int NTH = TEAMS * 96;
  static constexpr int ROW_BYTES = LW * 4;
  static constexpr int RING_BYTES = NSLOT * CH * ROW_BYTES;             // per team: 14 KB
  static constexpr int IN_W = 144;              // frame box: columns (xs-8) & ~15 .. + 143, PU rows
  static constexpr int NIN = 6;                 // frame boxes in flight per team
  static constexpr int IN_STAGE = (IN_W * PU + 127) / 128 * 128;        // 768 B
  static constexpr int OFF_IN = TEAMS * RING_BYTES;
  static constexpr int OFF_BAR = OFF_IN + TEAMS * NIN * IN_STAGE;
  static constexpr int NBAR = 2 * NSLOT + NIN;  // per team: full[NSLOT], empty[NSLOT], infull[NIN]
  static constexpr int SMEM = OFF_BAR + TEAMS * NBAR * 8;
};

// shared-memory accesses by 32-bit shared address (no generic -> shared conversion per access)
__device__ __forceinline__ void mbar_arrive_s(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(unsigned bar, unsigned phase) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "KLT_WAIT_S:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra KLT_DONE_S;\n\t"
      "bra KLT_WAIT_S;\n\t"
      "KLT_DONE_S:\n\t"
      "}" ::"r"(bar), "r"(phase)
      : "memory");
}
__device__ __forceinline__ void sts128_s(unsigned p, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128_s(unsigned p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(p));
  return v;
}
__device__ __forceinline__ uint2 lds64_s(unsigned p) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(p));
  return v;
}
__device__ __forceinline__ void mbar_expect_tx_s(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// (the fence orders the warp's generic-proxy reads of the stage before the async-proxy write)
__device__ __forceinline__ void tma_load_2d_s(unsigned dst, const CUtensorMap* map, int cx, int cy, unsigned bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<unsigned long long>(map)), "r"(cx), "r"(cy), "r"(bar)
      : "memory");
}
__device__ __forceinline__ uint2 ldg_u64(const unsigned char* p) {
  uint2 v;
  asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
#ifndef MARCH_L2_PREFETCH
#define MARCH_L2_PREFETCH 1
#endif
#ifndef MARCH_L0_BY_GX
#define MARCH_L0_BY_GX 1
#endif
#ifndef MARCH_HINT_STORES
#define MARCH_HINT_STORES 1
#endif
__device__ __forceinline__ void stg128_m(float* p, const float4& v, unsigned long long pol) {
#if MARCH_HINT_STORES
  stg128(p, v, pol);
#else
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
#endif
}
__device__ __forceinline__ void march_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// byte b of a word as float: one I2F.U8 with a byte selector (conversion pipe: idle otherwise)
__device__ __forceinline__ float u8f(unsigned w, int b) { return (float)((w >> (8 * b)) & 0xffu); }
// first contribution of a running sum: 0 + v * k (the reference starts every sum at 0.0f)
__device__ __forceinline__ void fma4_first(float4& a, const float4& v, const TapsF& t, int m, bool exact) {
  a = make_float4(0.f, 0.f, 0.f, 0.f);
  fma4(a, v, t, m, exact);
}
// zero bands of a completed row (columns c .. c+3 of row y, radius R), border tasks only
__device__ __forceinline__ float4 march_zero_bands(float4 a, int c, int y, int W, int H, int R) {
  if (y < R || y >= H - R) return make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < R || c >= W - R) a.x = 0.f;
  if (c + 1 < R || c + 1 >= W - R) a.y = 0.f;
  if (c + 2 < R || c + 2 >= W - R) a.z = 0.f;
  if (c + 3 < R || c + 3 >= W - R) a.w = 0.f;
  return a;
}

struct MarchTask { int xs, ya, yb, nring; bool border; };
__device__ __forceinline__ MarchTask march_task(int task, int nstrips, int seg_rows, int W, int H) {
  MarchTask k;
  const int seg = task / nstrips;
  k.xs = (task - seg * nstrips) * MarchGeo::SWI;
  k.ya = seg * seg_rows;
  k.yb = min(k.ya + seg_rows, H);
  // L0 rows that go through the ring: ya-3 .. yb+2, padded to whole chunks (the padding rows are
  // computed from clamped frame rows and stored nowhere)
  k.nring = (k.yb - k.ya + 2 * MarchGeo::RG + MarchGeo::CH - 1) / MarchGeo::CH * MarchGeo::CH;
  k.border = k.xs < 8 || k.xs + MarchGeo::LW > W || k.ya < 5 || k.ya + k.nring + 2 > H;
  return k;
}

// ---- producer: frame rows ya-5 .. -> L0 rows ya-3 .. (ring, nring of them) and ya .. yb-1 (HBM) -----
// The frame rows arrive through TMA: boxes of 144 bytes x 5 rows (one group of the unrolled loop),
// NIN groups in flight, issued by lane 0 as soon as the warp has consumed the box that occupied the
// stage.  (Direct loads, even issued 5 rows ahead into a register ring and backed by an L2 prefetch of
// the whole segment, left the producer on their scoreboard 29 % of its time: under this kernel's write
// load a DRAM read takes microseconds.  Without the L2 prefetch: 47 us instead of 35 us.)
template <bool EXACT>
__device__ __forceinline__ void march_produce(const CUtensorMap* map, int W, int H,
                                              const MarchTask k, const TapsF& ts, float* __restrict__ out_img,
                                              int opitch, unsigned ring_s, unsigned full_s, unsigned empty_s,
                                              unsigned in_s, unsigned infull_s, unsigned& chunk, unsigned& ingroup,
                                              bool& waited) {
  using G = MarchGeo;
  constexpr int RS = G::RS;
  const int lane = threadIdx.x & 31;
  const bool odd = lane & 1;
  const int c = k.xs - 4 + 4 * lane;                       // my four columns: c .. c+3
  // even lanes read columns c-4 .. c+3 (left neighbour word + own), odd lanes c .. c+7 (own + right
  // neighbour word): both 8-byte aligned; the missing neighbour is the partner lane's own word
  const int xbox = (k.xs - 8) & ~15;                       // box origin: 16-byte aligned (TMA), <= xs-8
  const unsigned lo = (unsigned)((odd ? c : c - 4) - xbox);   // my 8 bytes inside a box row: 0 .. 136
  const int y0 = k.ya - (RS + G::RG);                      // frame row of step 0
  const int NR = k.nring + 2 * RS;                         // frame rows walked
  const int ngroups = (NR + G::PU - 1) / G::PU;
  if (lane == 0) {
    const int npro = ngroups < G::NIN ? ngroups : G::NIN;
    for (int g = 0; g < npro; ++g) {
      const unsigned st = (ingroup + g) % G::NIN;
      mbar_expect_tx_s(infull_s + 8 * st, G::IN_W * G::PU);
      tma_load_2d_s(in_s + st * G::IN_STAGE, map, xbox, y0 + G::PU * g, infull_s + 8 * st);
    }
  }
  // Nothing read so far comes from the predecessor kernel; the pyramid slot written below may still be
  // read by it (the previous frame's tracker): wait in front of the first store.
  if (!waited) { pdl_wait(); waited = true; }
#if !MARCH_L0_BY_GX
  const bool st_lane = lane >= 1 && lane <= 30 && c < W;
  const unsigned long long pol = store_policy(true);
  float* outp = out_img + (size_t)(y0 - RS) * opitch + c;  // L0 row completed by step i
  const int st0 = k.ya - (y0 - RS), st1 = k.yb - (y0 - RS);   // steps whose completed row is stored to HBM
#endif
  float4 acc[G::PU];
#pragma unroll
  for (int u = 0; u < G::PU; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  unsigned slot = chunk % G::NSLOT;
  unsigned rp = ring_s + slot * (G::CH * G::ROW_BYTES) + 16 * lane;   // ring row of the next completed sum
  int rin = 0;
  for (int t = 0; t < ngroups; ++t) {
    const unsigned st = ingroup % G::NIN;
    mbar_wait_s(infull_s + 8 * st, (ingroup / G::NIN) & 1);
    const unsigned ip = in_s + st * G::IN_STAGE + lo;
#pragma unroll
    for (int u = 0; u < G::PU; ++u) {
      const int i = G::PU * t + u;
      if (i < NR) {
        const uint2 w = lds64_s(ip + u * G::IN_W);
        const unsigned sh = __shfl_xor_sync(0xffffffffu, odd ? w.x : w.y, 1);
        const unsigned wl = odd ? sh : w.x, cur = odd ? w.x : w.y, wr = odd ? w.y : sh;
        float px[8];                                       // columns c-2 .. c+5
        px[0] = u8f(wl, 2); px[1] = u8f(wl, 3);
        px[2] = u8f(cur, 0); px[3] = u8f(cur, 1); px[4] = u8f(cur, 2); px[5] = u8f(cur, 3);
        px[6] = u8f(wr, 0); px[7] = u8f(wr, 1);
        float hs[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float a = 0.0f;
#pragma unroll
          for (int m = 0; m < 2 * RS + 1; ++m) a = mac<EXACT>(a, px[q + m], ts.k[m]);
          hs[q] = a;
        }
        const float4 v = make_float4(hs[0], hs[1], hs[2], hs[3]);
        // frame row i is tap m of the L0 row that started m rows ago (running sum (u - m) mod 5)
        fma4_first(acc[u], v, ts, 0, EXACT);
#pragma unroll
        for (int m = 1; m < 2 * RS + 1; ++m) fma4(acc[(u - m + G::PU) % G::PU], v, ts, m, EXACT);
        if (i >= 2 * RS) {                                 // running sum (u + 1) mod 5 is complete
          const int j = (u + 1) % G::PU;
          if (k.border) {                                  // (__syncwarp: keeps the block a branch, not predicated code)
            __syncwarp();
            acc[j] = march_zero_bands(acc[j], c, y0 - RS + i, W, H, RS);
          }
          if (rin == 0 && chunk >= G::NSLOT) mbar_wait_s(empty_s + 8 * slot, ((chunk / G::NSLOT) - 1) & 1);
          sts128_s(rp, acc[j]);
#if !MARCH_L0_BY_GX
          if (st_lane && (unsigned)(i - st0) < (unsigned)(st1 - st0)) stg128_m(outp, acc[j], pol);
#endif
          rp += G::ROW_BYTES;
          if (++rin == G::CH) {
            __syncwarp();
            if (lane == 0) mbar_arrive_s(full_s + 8 * slot);
            ++chunk;
            slot = chunk % G::NSLOT;
            rp = ring_s + slot * (G::CH * G::ROW_BYTES) + 16 * lane;
            rin = 0;
          }
        }
#if !MARCH_L0_BY_GX
        outp += opitch;
#endif
      }
    }
    __syncwarp();                                          // every lane has read the box: its stage is free
    if (lane == 0 && t + G::NIN < ngroups) {
      mbar_expect_tx_s(infull_s + 8 * st, G::IN_W * G::PU);
      tma_load_2d_s(in_s + st * G::IN_STAGE, map, xbox, y0 + G::PU * (t + G::NIN), infull_s + 8 * st);
    }
    ++ingroup;
  }
}

// ---- consumer: nring L0 rows of the ring (ya-3 ..) -> gx (GY = false) or gy (GY = true) rows ya .. yb-1 ----
// th: horizontal taps (derivative for gx, Gaussian for gy), tv: vertical taps (Gaussian for gx,
// derivative for gy).  The derivative's centre tap is exactly 0 and skipped (klt_fused.cuh).
template <bool EXACT, bool GY>
__device__ __forceinline__ void march_consume(unsigned ring_s, unsigned full_s, unsigned empty_s, unsigned& chunk,
                                              const MarchTask k, int W, int H, const TapsF& th, const TapsF& tv,
                                              float* __restrict__ out, int opitch, float* __restrict__ out_img) {
  using G = MarchGeo;
  constexpr int RG = G::RG, NT = 2 * RG + 1;
  const int lane = threadIdx.x & 31;
  const int li = lane < 30 ? lane : 29;                    // lanes 30, 31 shadow lane 29 and store nothing
  const int c = k.xs + 4 * li;
  const bool st_lane = lane < 30 && c < W;
  const unsigned long long pol = store_policy(false);
  float4 acc[NT];
#pragma unroll
  for (int u = 0; u < NT; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  float* outp = out + (size_t)(k.ya - 2 * RG) * opitch + c;      // output row completed by ring row i
  const int nvalid = k.yb - k.ya;                          // ring rows 6 .. 6 + nvalid - 1 complete a stored row
#if MARCH_L0_BY_GX
  float* outl = out_img + (size_t)(k.ya - RG) * opitch + c;      // L0 row held by ring row i
  const unsigned long long pol_img = store_policy(true);
#endif
  for (int i0 = 0; i0 < k.nring; i0 += NT) {
    const unsigned slot = chunk % G::NSLOT;
    mbar_wait_s(full_s + 8 * slot, (chunk / G::NSLOT) & 1);
    const unsigned p = ring_s + slot * (G::CH * G::ROW_BYTES) + 16 * li;   // ring column 0 <-> xs-4: my window starts at c-4
#pragma unroll
    for (int u = 0; u < NT; ++u) {
      const float4 a0 = lds128_s(p + u * G::ROW_BYTES), a1 = lds128_s(p + u * G::ROW_BYTES + 16),
                   a2 = lds128_s(p + u * G::ROW_BYTES + 32);
      const float win[12] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
#if MARCH_L0_BY_GX
      if (!GY) {                                           // the gx warp also writes L0 (its middle float4) to HBM
        if (st_lane && (unsigned)(i0 + u - RG) < (unsigned)nvalid) stg128_m(outl, a1, pol_img);
        outl += opitch;
      }
#endif
      float h[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float a = 0.0f;
#pragma unroll
        for (int m = 0; m < NT; ++m)
          if (GY || m != RG) a = mac<EXACT>(a, win[q + 1 + m], th.k[m]);
        h[q] = a;
      }
      const float4 v = make_float4(h[0], h[1], h[2], h[3]);
      fma4_first(acc[u], v, tv, 0, EXACT);
#pragma unroll
      for (int m = 1; m < NT; ++m)
        if (!GY || m != RG) fma4(acc[(u - m + NT) % NT], v, tv, m, EXACT);
      const int j = (u + 1) % NT;                          // the running sum that is complete now
      if (k.border) {
        __syncwarp();
        acc[j] = march_zero_bands(acc[j], c, k.ya - 2 * RG + i0 + u, W, H, RG);
      }
      if (st_lane && (unsigned)(i0 + u - 2 * RG) < (unsigned)nvalid) stg128_m(outp, acc[j], pol);
      outp += opitch;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_s(empty_s + 8 * slot);
    ++chunk;
  }
}

// Tasks [task0, ntasks) = (segment, strip) pairs in raster order, dealt round-robin to the teams of
// the grid (whole segments per launch when a frame is built band by band behind its upload).
template <bool EXACT>
__global__ void __launch_bounds__(MarchGeo::NTH, 2)
l0_march_kernel(const __grid_constant__ CUtensorMap map, int W, int H, int nstrips, int seg_rows,
                int task0, int ntasks, TapsF ts, TapsF tg, TapsF td, float* __restrict__ out_img,
                float* __restrict__ out_gx, float* __restrict__ out_gy, int opitch) {
  using G = MarchGeo;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int team = warp / 3, role = warp - 3 * team;
  const unsigned ring_s = smem_u32(smem_raw) + team * G::RING_BYTES;
  const unsigned in_s = smem_u32(smem_raw) + G::OFF_IN + team * (G::NIN * G::IN_STAGE);
  const unsigned full_s = smem_u32(smem_raw) + G::OFF_BAR + team * (G::NBAR * 8);
  const unsigned empty_s = full_s + G::NSLOT * 8;
  const unsigned infull_s = empty_s + G::NSLOT * 8;
  if (threadIdx.x < G::TEAMS * G::NBAR) {
    unsigned long long* b = reinterpret_cast<unsigned long long*>(smem_raw + G::OFF_BAR) + threadIdx.x;
    const int which = threadIdx.x % G::NBAR;
    mbar_init(b, (which >= G::NSLOT && which < 2 * G::NSLOT) ? 2u : 1u);   // empty: gx + gy arrive; full / infull: one
  }
  if (threadIdx.x == 0)
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<unsigned long long>(&map)) : "memory");
  __syncthreads();
  unsigned chunk = 0, ingroup = 0;
  bool waited = false;
  const int stride = gridDim.x * G::TEAMS;
  for (int task = task0 + blockIdx.x * G::TEAMS + team; task < ntasks; task += stride) {
    const MarchTask k = march_task(task, nstrips, seg_rows, W, H);
    if (role == 0) {
      march_produce<EXACT>(&map, W, H, k, ts, out_img, opitch, ring_s, full_s, empty_s, in_s, infull_s, chunk, ingroup, waited);
    } else {
      if (!waited) { pdl_wait(); waited = true; }
      if (role == 1) march_consume<EXACT, false>(ring_s, full_s, empty_s, chunk, k, W, H, td, tg, out_gx, opitch, out_img);
      else march_consume<EXACT, true>(ring_s, full_s, empty_s, chunk, k, W, H, tg, td, out_gy, opitch, out_img);
    }
  }
  if (!waited) pdl_wait();
  pdl_launch_dependents();
}
