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
//   producer  lane = 4 columns (strip + 4 halo columns each side).  Per frame row: one 64-bit shared load
//             from the TMA box (its own 4 pixels + one neighbour word), the other neighbour's pixels by
//             one butterfly shuffle of the raw word, horizontal 5-tap Gaussian in registers, then the row
//             is SCATTERED into the five running vertical sums it belongs to (taps in increasing row
//             order == the reference's summation order).  The sum that completes is a row of L0: it
//             goes into a 35-row ring in shared memory.
//   gx, gy    two consumer warps, lane = 4 columns.  Per L0 row of the ring: three 128-bit shared loads
//             (12 columns), horizontal derivative (gx) or Gaussian (gy) 7-tap in registers, scatter into
//             seven running vertical sums, the completed row goes to HBM; the gx warp also stores L0.
// Vertical passes therefore need no halo recomputation inside a segment (only 10 / 6 warm-up rows at
// its top), no intermediate but the L0 ring touches shared memory, and the only synchronisation is a
// set of mbarriers on a 35-row ring between the three warps of a team (full, one per 5-row producer
// chunk; empty, one per 7-row consumer group).  Teams are independent: 4 per CTA, 2 CTAs per SM.
//
// Border semantics as everywhere (klt_dev.cu header): outputs in the zero bands are written as zeros
// by the tasks that touch the image border; what those tasks read outside the image is clamped to the
// nearest pixel / row and only ever feeds zeroed outputs.  A zero band of a horizontal pass is applied
// to the completed row of the vertical pass that follows it (a column of zeros sums to zero), so each
// warp has ONE fix-up per row, behind a warp-uniform branch that interior tasks never take.
//
// What the versions taught (ncu, 4K frame; the full log is in DESIGN.md 4): (1) separate interior /
// border instantiations of the three roles made 88 KB of code: 45 % of the warp samples were "no
// instruction", 63 us.  (2) A load issued into a register-ring slot whose old value is still live lands
// in a temporary and is MOVed into the slot at the end of the row -- the move waits for the load: the
// DRAM latency of every row exposed, 52 us.  (3) Under this kernel's write load a DRAM read takes
// microseconds: the frame rows come through TMA boxes issued six groups ahead.  (4) Then the kernel is
// issue bound on control code: rows are branch free now (a ring of 35 rows = 7 producer chunks = 5
// consumer groups, so that nothing straddles the wrap and every offset is a constant; tasks padded to
// whole revolutions; zero bands applied by a second store behind a warp-uniform branch instead of
// modifying the running sums): producer 45 and consumers ~70 instructions per row, 18.8 M in total,
// 32.5 us -- and 27.6 us with the global stores removed: at 24 warps per SM (80 registers: the running
// sums) the three warps' own latencies are not hidden; prefetching the next row's window spills.
#pragma once

struct MarchGeo {
  static constexpr int SWI = 120;               // interior columns of a strip (30 lanes x 4)
  static constexpr int LW = 128;                // L0 columns in the ring: xs-4 .. xs+123
  static constexpr int RS = 2, RG = FUSED_RG;   // smoothing / gradient radii
  static constexpr int PU = 2 * RS + 1;         // producer: 5 running sums, chunks of 5 ring rows
  static constexpr int CU = 2 * RG + 1;         // consumers: 7 running sums, groups of 7 ring rows
  static constexpr int RING_ROWS = PU * CU;     // 35: one revolution = 7 chunks = 5 groups, nothing straddles the wrap
#ifndef MARCH_TEAMS
#define MARCH_TEAMS 4
#define MARCH_CPS 2
#endif
  static constexpr int TEAMS = MARCH_TEAMS;     // teams per CTA
  static constexpr int CPS = MARCH_CPS;         // CTAs per SM the kernel is compiled for
  static constexpr int NTH = TEAMS * 96;
  static constexpr int ROW_BYTES = LW * 4;
  static constexpr int RING_BYTES = RING_ROWS * ROW_BYTES;              // per team: 17.5 KB
  static constexpr int IN_W = 144;              // frame box: columns (xs-8) & ~15 .. + 143, PU rows
  static constexpr int NIN = 6;                 // frame boxes in flight per team
  static constexpr int IN_STAGE = (IN_W * PU + 127) / 128 * 128;        // 768 B
  static constexpr int OFF_IN = TEAMS * RING_BYTES;
  static constexpr int OFF_BAR = OFF_IN + TEAMS * NIN * IN_STAGE;
  static constexpr int NBAR = CU + PU + NIN;    // per team: full[7] (one per chunk), empty[5] (one per group), infull[NIN]
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
  // L0 rows that go through the ring: ya-3 .. yb+2, padded to whole revolutions (the padding rows are
  // computed from frame rows that exist or are zero-filled by TMA and are stored nowhere)
  k.nring = (k.yb - k.ya + 2 * MarchGeo::RG + MarchGeo::RING_ROWS - 1) / MarchGeo::RING_ROWS * MarchGeo::RING_ROWS;
  k.border = k.xs < 8 || k.xs + MarchGeo::LW > W || k.ya < 5 || k.ya + k.nring + 2 > H;
  return k;
}

// one frame row of the producer: box row -> horizontal Gaussian -> scatter into the running sums;
// U = the row's position in the 5-cycle (frame row i = 5 t + U): it starts sum U and completes sum (U + 1) % 5
template <bool EXACT, int U>
__device__ __forceinline__ void march_prow(unsigned ip, bool odd, const TapsF& ts, float4 (&acc)[MarchGeo::PU]) {
  using G = MarchGeo;
  const uint2 w = lds64_s(ip);
  const unsigned sh = __shfl_xor_sync(0xffffffffu, odd ? w.x : w.y, 1);
  const unsigned wl = odd ? sh : w.x, cur = odd ? w.x : w.y, wr = odd ? w.y : sh;
  float px[8];                                             // columns c-2 .. c+5
  px[0] = u8f(wl, 2); px[1] = u8f(wl, 3);
  px[2] = u8f(cur, 0); px[3] = u8f(cur, 1); px[4] = u8f(cur, 2); px[5] = u8f(cur, 3);
  px[6] = u8f(wr, 0); px[7] = u8f(wr, 1);
  float hs[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float a = 0.0f;
#pragma unroll
    for (int m = 0; m < G::PU; ++m) a = mac<EXACT>(a, px[q + m], ts.k[m]);
    hs[q] = a;
  }
  const float4 v = make_float4(hs[0], hs[1], hs[2], hs[3]);
  fma4_first(acc[U], v, ts, 0, EXACT);
#pragma unroll
  for (int m = 1; m < G::PU; ++m) fma4(acc[(U - m + G::PU) % G::PU], v, ts, m, EXACT);
}

// ---- producer: frame rows ya-5 .. -> L0 rows ya-3 .. (ring, nring of them) -----------------------
// The frame rows arrive through TMA: boxes of 144 bytes x 5 rows, NIN in flight, issued by lane 0 as
// soon as the warp has consumed the box that occupied the stage.  Four warm-up rows (sums that
// complete nothing), then one chunk of five ring rows per loop iteration, branch free: the chunk's
// ring rows and the box rows are at fixed offsets, the wait for the consumers sits at its top, the
// signal at its bottom.
template <bool EXACT>
__device__ __forceinline__ void march_produce(const CUtensorMap* map, int W, int H, const MarchTask k, const TapsF& ts,
                                              unsigned ring_s, unsigned full_s, unsigned empty_s, unsigned in_s,
                                              unsigned infull_s, unsigned& rev, unsigned& ingroup, bool& waited) {
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
  const int nchunks = k.nring / G::PU;
  // box 0: frame rows y0-1 .. y0+3 (steps 0 .. 3 in its rows 1 .. 4); box 1 + j: the five rows of chunk j
  const int nbox = nchunks + 1;
  if (lane == 0) {
    const int npro = nbox < G::NIN ? nbox : G::NIN;
    for (int g = 0; g < npro; ++g) {
      const unsigned st = (ingroup + g) % G::NIN;
      mbar_expect_tx_s(infull_s + 8 * st, G::IN_W * G::PU);
      tma_load_2d_s(in_s + st * G::IN_STAGE, map, xbox, y0 - 1 + G::PU * g, infull_s + 8 * st);
    }
  }
  if (!waited) { pdl_wait(); waited = true; }              // (nothing above touches what the predecessor wrote)
  float4 acc[G::PU];
#pragma unroll
  for (int u = 0; u < G::PU; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  {                                                        // warm-up: steps 0 .. 3
    const unsigned st = ingroup % G::NIN;
    mbar_wait_s(infull_s + 8 * st, (ingroup / G::NIN) & 1);
    const unsigned ip = in_s + st * G::IN_STAGE + lo;
    march_prow<EXACT, 0>(ip + 1 * G::IN_W, odd, ts, acc);
    march_prow<EXACT, 1>(ip + 2 * G::IN_W, odd, ts, acc);
    march_prow<EXACT, 2>(ip + 3 * G::IN_W, odd, ts, acc);
    march_prow<EXACT, 3>(ip + 4 * G::IN_W, odd, ts, acc);
    __syncwarp();
    if (lane == 0 && G::NIN < nbox) {
      mbar_expect_tx_s(infull_s + 8 * st, G::IN_W * G::PU);
      tma_load_2d_s(in_s + st * G::IN_STAGE, map, xbox, y0 - 1 + G::PU * G::NIN, infull_s + 8 * st);
    }
    ++ingroup;
  }
  int cc = 0;                                              // chunk inside the revolution: 0 .. 6
  for (int j = 0; j < nchunks; ++j) {
    const unsigned st = ingroup % G::NIN;
    mbar_wait_s(infull_s + 8 * st, (ingroup / G::NIN) & 1);
    if (rev > 0) {                                         // the groups this chunk overlaps were read a revolution ago?
      const int e0 = (G::PU * cc) / G::CU, e1 = (G::PU * cc + G::PU - 1) / G::CU;
      mbar_wait_s(empty_s + 8 * e0, (rev - 1) & 1);
      if (e1 != e0) mbar_wait_s(empty_s + 8 * e1, (rev - 1) & 1);
    }
    const unsigned ip = in_s + st * G::IN_STAGE + lo;
    const unsigned rp = ring_s + cc * (G::PU * G::ROW_BYTES) + 16 * lane;
    // (a zero band rewrites the row in the ring right away: the running sum is restarted by the next row)
#define KLT_MARCH_PROW(U, RR)                                                                             \
    march_prow<EXACT, U>(ip + RR * G::IN_W, odd, ts, acc);                                                \
    sts128_s(rp + RR * G::ROW_BYTES, acc[RR]);                                                            \
    if (k.border) {                                                                                       \
      __syncwarp();                                        /* (keeps the block a branch, not predicated code) */ \
      sts128_s(rp + RR * G::ROW_BYTES, march_zero_bands(acc[RR], c, k.ya - G::RG + G::PU * j + RR, W, H, RS));   \
    }
    KLT_MARCH_PROW(4, 0) KLT_MARCH_PROW(0, 1) KLT_MARCH_PROW(1, 2) KLT_MARCH_PROW(2, 3) KLT_MARCH_PROW(3, 4)
#undef KLT_MARCH_PROW
    __syncwarp();                                          // every lane has read the box and written its rows
    if (lane == 0) {
      mbar_arrive_s(full_s + 8 * cc);
      if (j + 1 + G::NIN < nbox) {
        mbar_expect_tx_s(infull_s + 8 * st, G::IN_W * G::PU);
        tma_load_2d_s(in_s + st * G::IN_STAGE, map, xbox, y0 - 1 + G::PU * (j + 1 + G::NIN), infull_s + 8 * st);
      }
    }
    ++ingroup;
    if (++cc == G::CU) { cc = 0; ++rev; }
  }
}

// ---- consumer: nring L0 rows of the ring (ya-3 ..) -> gx (GY = false) or gy (GY = true) rows ya .. yb-1 ----
// th: horizontal taps (derivative for gx, Gaussian for gy), tv: vertical taps (Gaussian for gx,
// derivative for gy).  The derivative's centre tap is exactly 0 and skipped (klt_fused.cuh).  One group
// of seven ring rows per loop iteration at fixed offsets; the gx warp also stores L0 (the middle float4
// of its window) -- the producer is each team's critical path.
template <bool EXACT, bool GY>
__device__ __forceinline__ void march_consume(unsigned ring_s, unsigned full_s, unsigned empty_s, unsigned& rev,
                                              const MarchTask k, int W, int H, const TapsF& th, const TapsF& tv,
                                              float* __restrict__ out, int opitch, float* __restrict__ out_img) {
  using G = MarchGeo;
  constexpr int RG = G::RG, NT = G::CU;
  const int lane = threadIdx.x & 31;
  const int li = lane < 30 ? lane : 29;                    // lanes 30, 31 shadow lane 29 and store nothing
  const int c = k.xs + 4 * li;
#ifdef MARCH_NO_STORES
  const bool st_lane = lane < 30 && c < W && opitch < 0;     // experiment: everything but the global stores
#else
  const bool st_lane = lane < 30 && c < W;
#endif
  const unsigned long long pol = store_policy(false);
  float4 acc[NT];
#pragma unroll
  for (int u = 0; u < NT; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  float* outp = out + (size_t)(k.ya - 2 * RG) * opitch + c;      // output row completed by ring row i
  float* outl = out_img + (size_t)(k.ya - RG) * opitch + c;      // L0 row held by ring row i
  const unsigned long long pol_img = store_policy(true);
  const int nvalid = k.yb - k.ya;                          // ring rows 6 .. 6 + nvalid - 1 complete a stored row
  int gg = 0;                                              // group inside the revolution: 0 .. 4
  for (int i0 = 0; i0 < k.nring; i0 += NT) {
    {                                                      // the chunks this group overlaps are written?
      const int f0 = (NT * gg) / G::PU, f1 = (NT * gg + NT - 1) / G::PU;
      for (int f = f0; f <= f1; ++f) mbar_wait_s(full_s + 8 * f, rev & 1);
    }
    const unsigned p = ring_s + gg * (NT * G::ROW_BYTES) + 16 * li;   // ring column 0 <-> xs-4: my window starts at c-4
#pragma unroll
    for (int u = 0; u < NT; ++u) {
      const float4 a0 = lds128_s(p + u * G::ROW_BYTES), a1 = lds128_s(p + u * G::ROW_BYTES + 16),
                   a2 = lds128_s(p + u * G::ROW_BYTES + 32);
      const float win[12] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
      if (!GY) {                                           // L0 goes to HBM from here
        if (st_lane && (unsigned)(i0 + u - RG) < (unsigned)nvalid) stg128(outl, a1, pol_img);
        outl += opitch;
      }
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
      const bool st_row = st_lane && (unsigned)(i0 + u - 2 * RG) < (unsigned)nvalid;
      if (st_row) stg128(outp, acc[j], pol);
      if (k.border) {                                      // zero bands: store the row again (same thread: ordered)
        __syncwarp();
        if (st_row) stg128(outp, march_zero_bands(acc[j], c, k.ya - 2 * RG + i0 + u, W, H, RG), pol);
      }
      outp += opitch;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive_s(empty_s + 8 * gg);
    if (++gg == G::PU) { gg = 0; ++rev; }
  }
}

// Tasks [task0, ntasks) = (segment, strip) pairs in raster order, dealt round-robin to the teams of
// the grid (whole segments per launch when a frame is built band by band behind its upload).
template <bool EXACT>
__global__ void __launch_bounds__(MarchGeo::NTH, MarchGeo::CPS)
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
  const unsigned empty_s = full_s + G::CU * 8;
  const unsigned infull_s = empty_s + G::PU * 8;
  if (threadIdx.x < G::TEAMS * G::NBAR) {
    unsigned long long* b = reinterpret_cast<unsigned long long*>(smem_raw + G::OFF_BAR) + threadIdx.x;
    const int which = threadIdx.x % G::NBAR;
    mbar_init(b, (which >= G::CU && which < G::CU + G::PU) ? 2u : 1u);   // empty: gx + gy arrive; full / infull: one
  }
  if (threadIdx.x == 0)
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<unsigned long long>(&map)) : "memory");
  __syncthreads();
  unsigned rev = 0, ingroup = 0;                           // ring revolutions completed / frame boxes consumed
  bool waited = false;
  const int stride = gridDim.x * G::TEAMS;
  for (int task = task0 + blockIdx.x * G::TEAMS + team; task < ntasks; task += stride) {
    const MarchTask k = march_task(task, nstrips, seg_rows, W, H);
    if (role == 0) {
      march_produce<EXACT>(&map, W, H, k, ts, ring_s, full_s, empty_s, in_s, infull_s, rev, ingroup, waited);
    } else {
      if (!waited) { pdl_wait(); waited = true; }
      if (role == 1) march_consume<EXACT, false>(ring_s, full_s, empty_s, rev, k, W, H, td, tg, out_gx, opitch, out_img);
      else march_consume<EXACT, true>(ring_s, full_s, empty_s, rev, k, W, H, tg, td, out_gy, opitch, out_img);
    }
  }
  if (!waited) pdl_wait();
  pdl_launch_dependents();
}
