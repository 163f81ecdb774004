// klt_mega.cuh -- the whole pyramid of one frame in ONE persistent launch (included by klt_dev.cu).
//
// pyramid_mega_kernel: u8 frame -> L_l, gx_l, gy_l for every level l.  Replaces the chain
// l0_fused_kernel + (nlevels-1) x level_fused_kernel (reference: the pyramid part of
// KLTTrackFeatures, src/V1/trackFeatures.c:1309-1321 = convolve.c:37-314 + pyramid.c:87-131)
// with a tile-granular dataflow schedule:
//
//   * work items = tiles of every level (64x64 at level 0, TXxTY above) in two queues, each an
//     atomic counter: level-0 tiles in raster order, and the tiles of the coarser levels in a
//     topological order listed by the host (klt_dev.cu: mega_schedule): a tile row of level l
//     appears after the tile rows of level l-1 whose pixels it reads.  Two CTAs in three serve the
//     level-0 queue, the third the coarse queue (tile types then rarely alternate inside a CTA: a
//     coarse source box cannot be prefetched under a live level-0 tile); a coarse CTA whose next
//     item is not ready yet helps with level-0 tiles meanwhile, and level-0 CTAs move to the
//     coarse queue when theirs is empty.
//   * a tile of level l >= 1 waits until the tile rows of level l-1 under its source box are
//     complete (per-tile-row completion counters in global memory, release/acquire); a level-0
//     tile waits until the u8 rows under its box have arrived from the host (a device word the
//     copy stream advances after every uploaded band, cuStreamWriteValue32) -- so the kernel is
//     launched once, before the frame is even on the device, and works its way down the frame and
//     up the pyramid behind the PCIe transfer.
//   * deadlock free: level-0 tiles depend on nothing but the upload; a coarse item depends on
//     level-0 tiles and on earlier coarse items, which were claimed earlier by CTAs that are
//     resident (grid <= resident capacity), signal completion before they wait again, and only
//     ever wait for strictly earlier items.
//
// Why one launch: measured on B200 (tools/overlap_probe.cu), a dependent kernel boundary costs
// 2.6 us on an idle bus but 7-12 us while an H2D copy is in flight (the front end fetches its
// commands over the same PCIe link), and every launch of a fused kernel over a partial frame has
// ~8 us of fixed latency (cold I-cache, TMA descriptor fetch, one tile deep pipeline).  Four
// launches per band made the banded upload a loss; the tile queue makes it a win, and it also
// removes the three launch gaps and the small-grid tails of the coarse levels in the resident path.
//
// Tile code is the same as in klt_fused.cuh (l0_fused_tile*, lv_stage_*): results are bit-identical
// to the per-level kernels in both arithmetic modes.
#pragma once

static constexpr int MEGA_MAX_LEVELS = 8;
static constexpr int MEGA_MAX_SEGS = 1024;

struct MegaLevel {
  int W, H, pitch;               // level size and row pitch (floats)
  float *img, *gx, *gy;
  int tiles_x, tiles_y;
  int done_off;                  // tile row jr of this level signals done[done_off + jr]
  unsigned target;               // value of that counter once the row is complete in this frame
};
struct MegaSeg {                 // coarse work items [w0, w0 + n): tile rows jr0.. of `level`, row-major
  int level, jr0, w0, n;
};
struct MegaParams {
  CUtensorMap map[MEGA_MAX_LEVELS];   // source of level l: u8 frame (l = 0) or L_{l-1}
  MegaLevel lv[MEGA_MAX_LEVELS];
  int nlev, nseg;
  int nitems0, nitems1;               // level-0 tiles (row-major) / coarse items (segs order)
  const MegaSeg* segs;
  unsigned* ctr;                      // [0] level-0 queue, [1] coarse queue, [2] CTAs that left (the last resets all three)
  unsigned* done;                     // per tile row completion counters (never reset: see target)
  const unsigned* u8_flag;            // rows of the u8 frame on the device: *u8_flag - u8_base
  unsigned u8_base; int has_feed;
  int ready_below;                    // levels < ready_below were complete before the launch (tail mode)
  int serial;                         // debug: no prefetch, tile k is loaded after tile k-1 is finished
  int coarse_every;                   // CTA b starts on the coarse queue iff b % coarse_every == coarse_every - 1
  TapsF ts, tp, tg, td;
};

__device__ __forceinline__ unsigned ld_relaxed_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int SS, int R, int TX, int TY>
struct MegaGeo {
  using GL = LvGeo<SS, R, TX, TY>;
  static constexpr int SMEM_TILE = L0Geo::SMEM > GL::SMEM ? L0Geo::SMEM : GL::SMEM;
  static constexpr int OFF_BAR = (SMEM_TILE + 15) / 16 * 16;
  static constexpr int OFF_INFO = OFF_BAR + 32;        // 2 x 4 ints: level (-1 = no work left), x0, y0
  static constexpr int SMEM = OFF_INFO + 32;
  static constexpr int CPS = SMEM <= 75 * 1024 ? 3 : (SMEM <= 113 * 1024 ? 2 : 1);
};

// dependencies of the tile at (level, y0): true if they are satisfied now
template <int SS, int R, int TX, int TY>
__device__ __forceinline__ bool mega_deps_ready(const MegaParams& P, int level, int y0) {
  if (level == 0) {
    if (!P.has_feed) return true;
    int need = y0 + L0Geo::TY + L0Geo::RS + L0Geo::RG;
    if (need > P.lv[0].H) need = P.lv[0].H;
    // written by the copy engine in stream order behind the band it announces; both land in L2 /
    // HBM, which is where this load and the TMA read go: a relaxed system-scope load is enough
    // (an acquire.sys poll from every CTA slowed the upload itself down, measured)
    return (int)(ld_relaxed_sys_u32(P.u8_flag) - P.u8_base) >= need;
  }
  using GL = LvGeo<SS, R, TX, TY>;
  if (level - 1 < P.ready_below) return true;
  const MegaLevel& s = P.lv[level - 1];
  int ylo = SS * y0 + GL::YOFF, yhi = ylo + GL::SH - 1;          // source rows under the TMA box
  if (ylo < 0) ylo = 0;
  if (yhi > s.H - 1) yhi = s.H - 1;
  const int tys = level == 1 ? L0Geo::TY : TY;
  for (int r = ylo / tys; r <= yhi / tys; ++r)
    if ((int)(ld_acquire_u32(P.done + s.done_off + r) - s.target) < 0) return false;
  return true;
}

template <int SS, int R, int TX, int TY>
__device__ __forceinline__ void mega_issue_tma(const MegaParams& P, unsigned char* smem, int level, int x0,
                                               int y0, unsigned long long* bar) {
  using GL = LvGeo<SS, R, TX, TY>;
  asm volatile("fence.proxy.async;" ::: "memory");     // acquired generic-proxy writes -> async-proxy reads
  if (level == 0) {
    mbar_expect_tx(bar, L0Geo::U8_W * L0Geo::U8_H);
    tma_load_2d(smem + L0Geo::OFF_U8, &P.map[0], x0 - 16, y0 - (L0Geo::RS + L0Geo::RG), bar);
  } else {
    mbar_expect_tx(bar, GL::SW * GL::SH * 4);
    tma_load_2d(smem + GL::OFF_SRC, &P.map[level], SS * x0 + GL::XOFF, SS * y0 + GL::YOFF, bar);
  }
}

// mbarrier helpers of the producer / consumer hand-over
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned long long* bar, unsigned phase) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
  return ok != 0;
}

// Scheduler state of one CTA (registers of the scheduling thread).
//   on_coarse : the CTA serves the coarse queue (initially one CTA in coarse_every; every CTA once
//               the level-0 queue is empty)
//   held      : a claimed coarse item whose dependencies were not ready yet (h_level < 0: none)
//   p_ctr / p_next : completion counters of the (at most two) tiles in flight whose completion has
//               not been published yet, oldest first (nullptr: none)
struct MegaSched {
  bool on_coarse, l0_empty, coarse_empty;
  int seg_cursor;
  int h_level, h_x0, h_y0;
  unsigned* p_ctr; unsigned* p_next; unsigned n_published;
};

template <int SS, int R, int TX, int TY>
__device__ __forceinline__ bool mega_claim_l0(const MegaParams& P, MegaSched& st, int& x0, int& y0) {
  if (st.l0_empty) return false;
  const int w = (int)atomicAdd(P.ctr + 0, 1u);
  if (w >= P.nitems0) { st.l0_empty = true; return false; }
  const int tx = P.lv[0].tiles_x;
  y0 = (w / tx) * L0Geo::TY;
  x0 = (w - (w / tx) * tx) * L0Geo::TX;
  return true;
}
template <int SS, int R, int TX, int TY>
__device__ __forceinline__ bool mega_claim_coarse(const MegaParams& P, MegaSched& st) {
  if (st.coarse_empty) return false;
  const int w = (int)atomicAdd(P.ctr + 1, 1u);
  if (w >= P.nitems1) { st.coarse_empty = true; return false; }
  int c = st.seg_cursor;
  while (c + 1 < P.nseg && P.segs[c + 1].w0 <= w) ++c;
  st.seg_cursor = c;
  const MegaSeg sg = P.segs[c];
  const int tx = P.lv[sg.level].tiles_x;
  const int t = w - sg.w0;
  st.h_level = sg.level;
  st.h_y0 = (sg.jr0 + t / tx) * TY;
  st.h_x0 = (t - (t / tx) * tx) * TX;
  return true;
}
// would the next level-0 tile (not claimed yet) find its rows on the device?
template <int SS, int R, int TX, int TY>
__device__ __forceinline__ bool mega_peek_l0_ready(const MegaParams& P, const MegaSched& st) {
  if (st.l0_empty) return false;
  const int w = (int)ld_relaxed_u32(P.ctr + 0);
  if (w >= P.nitems0) return false;
  return mega_deps_ready<SS, R, TX, TY>(P, 0, (w / P.lv[0].tiles_x) * L0Geo::TY);
}
// publish the completion of the oldest unpublished tile once the compute warps have finished it.
// done_count (shared) = tiles this CTA has finished: a monotonic counter, so nothing is lost if two
// tiles finish between two calls (an mbarrier phase bit would alias).
__device__ __forceinline__ unsigned ld_acquire_cta_shared(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_cta_shared_inc(unsigned* p) {
  asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(smem_u32(p)) : "memory");
}
__device__ __forceinline__ void mega_service(MegaSched& st, const unsigned* done_count) {
  if (st.p_ctr != nullptr && (int)(ld_acquire_cta_shared(done_count) - st.n_published) > 0) {
    __threadfence();
    red_release_add(st.p_ctr, 1u);
    st.p_ctr = st.p_next;
    st.p_next = nullptr;
    st.n_published += 1;
  }
}

// Shared-memory control block behind the tile buffers:
//   bar_full  (1 arrival + TMA bytes): the source box of tile k has landed / "no more tiles"
//   bar_empty (1 arrival): the compute warps have consumed the source box of tile k
//   done_count: tiles whose every store the compute warps have issued
//   info[2][4]: level (-1: stop), x0, y0 of tile k in slot k & 1 (done_count lives in the last word)
//
// Warp specialisation: warps 0-7 (256 threads) run the tile code of klt_fused.cuh unchanged and
// never touch the queues; thread 0 of warp 8 is the scheduler: it claims items, polls their
// dependencies, issues the TMA load of tile k+1 as soon as tile k's source box is consumed, and
// publishes tile completions.  (With the claims, acquire loads and fences on thread 0 of the
// compute warps, ncu showed barrier + long-scoreboard stalls of 12.8 warps per issue: every tile
// waited for its own bookkeeping.)
#ifdef KLT_MEGA_DEBUG
#define MEGA_STUCK(site, cnt, ...)                                                        \
  do { if (++(cnt) > (1u << 22)) { printf("mega stuck site %d cta %d " __VA_ARGS__); __trap(); } } while (0)
#else
#define MEGA_STUCK(site, cnt, ...) do { } while (0)
#endif

template <int SS, int R, int TX, int TY, bool EXACT>
__global__ void __launch_bounds__(288, (MegaGeo<SS, R, TX, TY>::CPS))
pyramid_mega_kernel(const __grid_constant__ MegaParams P) {
  using MG = MegaGeo<SS, R, TX, TY>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned long long* bar_full = reinterpret_cast<unsigned long long*>(smem_raw + MG::OFF_BAR);
  unsigned long long* bar_empty = bar_full + 1;
  unsigned* bar_done = reinterpret_cast<unsigned*>(smem_raw + MG::OFF_INFO) + 7;   // done_count (see mega_service)
  volatile int* info = reinterpret_cast<volatile int*>(smem_raw + MG::OFF_INFO);
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_empty, 1);
    *bar_done = 0u;
  }
  __syncthreads();

  if (tid >= 256) {
    // ------------------------------- scheduler ----------------------------------------------
    if (tid != 256) return;
    MegaSched st;
    st.on_coarse = P.nitems1 > 0 && (int)(blockIdx.x % (unsigned)P.coarse_every) == P.coarse_every - 1;
    st.l0_empty = P.nitems0 == 0; st.coarse_empty = P.nitems1 == 0;
    if (st.l0_empty) st.on_coarse = true;
    st.seg_cursor = 0; st.h_level = -1; st.h_x0 = st.h_y0 = 0;
    st.p_ctr = nullptr; st.p_next = nullptr; st.n_published = 0;
    unsigned ph_empty = 0;
    int prev_level = -1;
    unsigned dbg = 0; (void)dbg;
    for (int k = 0;; ++k) {
      // ---- choose tile k ----
      int level = -1, x0 = 0, y0 = 0;
      unsigned ns = 32;
      dbg = 0;
      for (;;) {
        mega_service(st, bar_done);
        if (!st.on_coarse) {
          if (mega_claim_l0<SS, R, TX, TY>(P, st, x0, y0)) { level = 0; break; }
          st.on_coarse = true;
        }
        if (st.h_level < 0) mega_claim_coarse<SS, R, TX, TY>(P, st);
        if (st.h_level >= 0 && mega_deps_ready<SS, R, TX, TY>(P, st.h_level, st.h_y0)) {
          level = st.h_level; x0 = st.h_x0; y0 = st.h_y0; st.h_level = -1;
          break;
        }
        // nothing ready upstairs: help level 0 (only with a tile whose rows are on the device
        // while a coarse item is held, so the held item is not stuck behind the upload)
        if ((st.h_level < 0 || mega_peek_l0_ready<SS, R, TX, TY>(P, st)) &&
            mega_claim_l0<SS, R, TX, TY>(P, st, x0, y0)) { level = 0; break; }
        if (st.h_level < 0) break;                           // both queues are empty
        MEGA_STUCK(1, dbg, "k %d held L%d y0 %d l0_empty %d\n", 1, (int)blockIdx.x, k, st.h_level, st.h_y0, (int)st.l0_empty);
        __nanosleep(ns); if (ns < 1024) ns <<= 1;
      }
      if (level < 0) {                                       // tell the compute warps to stop
        // bar_full must be in its next phase first: an arrival while tile k-1's bytes are still
        // in flight would be a second arrival on that phase (count 1) -- a device exception
        if (k > 0)
          while (!mbar_test(bar_empty, ph_empty)) { mega_service(st, bar_done); MEGA_STUCK(7, dbg, "k %d\n", 7, (int)blockIdx.x, k); }
        info[4 * (k & 1)] = -1;
        mbar_arrive(bar_full);
        break;
      }
      ns = 32;
      while (!mega_deps_ready<SS, R, TX, TY>(P, level, y0)) {          // (level-0 tiles: the upload)
        mega_service(st, bar_done);
        MEGA_STUCK(2, dbg, "k %d L%d y0 %d\n", 2, (int)blockIdx.x, k, level, y0);
        __nanosleep(ns); if (ns < 2048) ns <<= 1;
      }
      if (k > 0) {
        // the source buffer is free once tile k-1 has consumed its box; a coarse box is larger
        // than the u8 box and overlaps level-0 buffers that stay live until tile k-1 is finished
        while (!mbar_test(bar_empty, ph_empty)) { mega_service(st, bar_done); MEGA_STUCK(3, dbg, "k %d\n", 3, (int)blockIdx.x, k); }
        ph_empty ^= 1;
        if ((prev_level == 0 && level != 0) || P.serial)
          while (st.p_ctr != nullptr) { mega_service(st, bar_done); MEGA_STUCK(4, dbg, "k %d\n", 4, (int)blockIdx.x, k); }
      }
      info[4 * (k & 1) + 1] = x0; info[4 * (k & 1) + 2] = y0; info[4 * (k & 1)] = level;
      mega_issue_tma<SS, R, TX, TY>(P, smem_raw, level, x0, y0, bar_full);
      // tiles k-1 (being computed) and k (being loaded) may both be unpublished; k-2 is finished
      // (its successor's source box has been consumed) and leaves the queue here at the latest
      while (st.p_next != nullptr) { mega_service(st, bar_done); MEGA_STUCK(5, dbg, "k %d\n", 5, (int)blockIdx.x, k); }
      unsigned* const ctr_k = P.done + P.lv[level].done_off + y0 / (level == 0 ? L0Geo::TY : TY);
      if (st.p_ctr == nullptr) st.p_ctr = ctr_k; else st.p_next = ctr_k;
      prev_level = level;
    }
    while (st.p_ctr != nullptr) { mega_service(st, bar_done); MEGA_STUCK(6, dbg, "end\n", 6, (int)blockIdx.x); }
    // the last CTA to leave resets the queues for the next launch (every CTA has stopped claiming)
    __threadfence();
    if (atomicAdd(P.ctr + 2, 1u) == gridDim.x - 1) {
      P.ctr[0] = 0; P.ctr[1] = 0; P.ctr[2] = 0;
      __threadfence();
    }
    return;
  }

  // --------------------------------- compute warps --------------------------------------------
  for (int k = 0;; ++k) {
    // Only thread 0 polls (with back-off); the other 255 threads block in the hardware barrier and
    // then pass the mbarrier test at once.  256 threads spinning on try_wait would take issue
    // slots and shared-memory bandwidth from the CTAs of the same SM that do have work -- with
    // tiles waiting on an upload or on other tiles that is the common case, not the exception.
    if (tid == 0) {
      unsigned ns = 20;
      while (!mbar_test(bar_full, (unsigned)(k & 1))) { __nanosleep(ns); if (ns < 320) ns <<= 1; }
    }
    tile_sync();
    mbar_wait(bar_full, (unsigned)(k & 1));
    const int level = info[4 * (k & 1)];
    if (level < 0) break;
    const int x0 = info[4 * (k & 1) + 1], y0 = info[4 * (k & 1) + 2];
    const MegaLevel& L = P.lv[level];
    if (level == 0) {
      // tiles whose 8-pixel margin stays inside the image never meet a zero band
      const bool border = (x0 < 8) || (y0 < 8) || (x0 + L0Geo::TX + 8 > L.W) || (y0 + L0Geo::TY + 8 > L.H);
      if (border) l0_fused_tile<EXACT, true>(smem_raw, L.W, P.ts, x0);
      else l0_fused_tile<EXACT, false>(smem_raw, L.W, P.ts, x0);
      tile_sync();                     // source box consumed
      if (tid == 0) mbar_arrive(bar_empty);
      if (border) l0_fused_tile_rest<EXACT, true>(smem_raw, L.W, L.H, P.ts, P.tg, P.td, L.img, L.gx, L.gy, L.pitch, x0, y0);
      else l0_fused_tile_rest<EXACT, false>(smem_raw, L.W, L.H, P.ts, P.tg, P.td, L.img, L.gx, L.gy, L.pitch, x0, y0);
    } else {
      const MegaLevel& S = P.lv[level - 1];
      const bool border = (x0 < 8) || (y0 < 8) || (x0 + TX + 16 > L.W) || (y0 + TY + 16 > L.H);
      if (border) lv_stage_p1<EXACT, true, SS, R, TX, TY>(smem_raw, P.tp, x0, S.W);
      else lv_stage_p1<EXACT, false, SS, R, TX, TY>(smem_raw, P.tp, x0, S.W);
      tile_sync();
      if (tid == 0) mbar_arrive(bar_empty);
      if (border)
        lv_stage_rest<EXACT, true, SS, R, TX, TY>(smem_raw, P.tp, P.tg, P.td, S.H, L.W, L.H, L.img, L.gx, L.gy, L.pitch, x0, y0);
      else
        lv_stage_rest<EXACT, false, SS, R, TX, TY>(smem_raw, P.tp, P.tg, P.td, S.H, L.W, L.H, L.img, L.gx, L.gy, L.pitch, x0, y0);
    }
    tile_sync();                       // all stores of this tile issued; shared buffers free
    if (tid == 0) red_release_cta_shared_inc(bar_done);
  }
}
