#!/usr/bin/env python
"""bench.py -- KLT hot-path throughput on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W          (this repo's CUDA path)
  python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU path)

Workload (BASELINE.json configs[3], the configuration the metric is quoted on):
synthetic 3840x2160 translated sequence, 4096 features, 4 pyramid levels
(subsampling 2), 7x7 window.  One step = one KLTTrackFeatures call: build the
three pyramids of one new frame, then track every live feature into it.
Metric: tracked features/s = features with val >= 0 entering the calls / time.
At N > 1 every rank runs its own independent sequence on its own GPU (no
collective on the data path); value is the sum over ranks / max time over ranks.

One JSON line on stdout (rank 0).  Keys are described in DESIGN.md section 6.
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "klt-feature-tracker-acceleration-gpus_b200"

WORKLOADS = {
    # name: (ncols, nrows, nfeatures, nlevels, subsampling, window, distinct frames)
    "4k": (3840, 2160, 4096, 4, 2, 7, 20),
    "1080p": (1920, 1080, 1024, 2, 4, 7, 64),
    "vga": (640, 480, 1000, 2, 4, 7, 64),
}
METRIC = "tracked_features_per_s"
UNIT = "features/s"


_JSON_FD = None          # the real stdout when fd 1 has been pointed at stderr (multi-rank runs)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def emit_json(obj):
    line = json.dumps(obj) + "\n"
    if _JSON_FD is None:
        sys.stdout.write(line)
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line.encode())


def algorithmic_bytes(ncols, nrows, nlevels, ss):
    """per new frame, read the u8 frame once + write the three f32
    pyramids: W*H + 12 * sum_l W_l*H_l (integer division per level)."""
    px, w, h = [], ncols, nrows
    for _ in range(nlevels):
        px.append(w * h)
        w //= ss
        h //= ss
    per_kernel = {
        "smooth_u8_tile": 5 * px[0],                                   # read u8, write L0
        "grad_tile": 12 * sum(px),                                     # read L_l, write gx_l, gy_l
        "pyrdown_tile": sum(4 * px[l - 1] + 4 * px[l] for l in range(1, nlevels)),
        "l0_fused_kernel": 13 * px[0],                                 # read u8, write L0, gx0, gy0
        "pyramid_mega_kernel": ncols * nrows + 12 * sum(px),           # whole pyramid in one launch (opt-in)
        # one pyramid step + that level's gradients: read L_{l-1}, write L_l, gx_l, gy_l
        "level_fused_kernel[level 1]": (4 * px[0] + 12 * px[1]) if nlevels > 1 else 0,
        "level_fused_kernel[level 2]": (4 * px[1] + 12 * px[2]) if nlevels > 2 else 0,
        "level_fused_kernel[level 3+]": sum(4 * px[l - 1] + 12 * px[l] for l in range(3, nlevels)),
        "frame_pipeline": ncols * nrows + 12 * sum(px),                # fully fused floor
    }
    # when level 0 is fused, the stand-alone gradient kernel only sees the coarser levels
    per_kernel["grad_tile_coarse"] = 12 * sum(px[1:])
    return per_kernel, px


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_frames(synth, ncols, nrows, nframes, seed, out):
    for t in range(nframes):
        synth.frame(ncols, nrows, seed=seed, t=float(t), out=out[t])


def setup_tc(L, nlevels, ss, window, device=None):
    tc = L.KLTCreateTrackingContext()
    t = tc.contents
    t.sequentialMode = 1
    t.window_width = t.window_height = window
    t.nPyramidLevels, t.subsampling = nlevels, ss
    if os.environ.get("KLT_BENCH_MAXIT"):          # diagnostic: how much of the tracker is its slowest features
        t.max_iterations = int(os.environ["KLT_BENCH_MAXIT"])
    L.KLTUpdateTCBorder(tc)
    if device is not None:
        L.KLTB200SetDevice(tc, device)
    return tc


# --------------------------------------------------------------------------- B200 arm
def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    pkg = importlib.import_module(PKG)
    rt = importlib.import_module(PKG + ".runtime")
    synth = importlib.import_module(PKG + ".synth")
    capi = pkg.capi
    L = rt.load()
    L.require_gpu()
    L.KLTSetVerbosity(0)
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL prints its version banner / debug lines on the process's stdout (fd 1, from C): point
        # fd 1 at stderr for the whole run and keep the real stdout for the one JSON line
        global _JSON_FD
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ncols, nrows, nfeat, nlevels, ss, window, nframes = WORKLOADS[args.workload]
    K, W = args.steps, args.warmup
    tc = setup_tc(L, nlevels, ss, window, device=local_rank)
    dev = L.KLTB200Device(tc)

    # ---- synthetic sequence: pinned host copy + HBM-resident copy -------------------
    t0 = time.time()
    # (KLT_BENCH_PAGEABLE=1: ordinary host memory, what a driver that mallocs its frames gets)
    frames_h = torch.empty((nframes, nrows, ncols), dtype=torch.uint8,
                           pin_memory=not os.environ.get("KLT_BENCH_PAGEABLE"))
    fh = frames_h.numpy()
    make_frames(synth, ncols, nrows, nframes, 12345 + rank, fh)
    frames_d = frames_h.cuda(non_blocking=False)
    torch.cuda.synchronize()
    fbytes = ncols * nrows
    d_ptr = lambda i: frames_d.data_ptr() + i * fbytes
    h_ptr = lambda i: frames_h.data_ptr() + i * fbytes
    log("[rank %d] %d frames %dx%d generated in %.1fs" % (rank, nframes, ncols, nrows, time.time() - t0))

    # ---- select on frame 0 (not part of the metric) ----------------------------------
    fl = L.KLTCreateFeatureList(nfeat)
    t0 = time.time()
    L.KLTSelectGoodFeatures(tc, C.c_void_p(h_ptr(0)), ncols, nrows, fl)
    t_select = time.time() - t0
    t0 = time.time()
    L.KLTSelectGoodFeatures(tc, C.c_void_p(h_ptr(0)), ncols, nrows, fl)
    t_select2 = time.time() - t0
    sel = capi.featurelist_to_arrays(fl)
    nsel = int((sel[2] > 0).sum())
    log("[rank %d] selected %d/%d features (first call %.1f ms, second %.1f ms)"
        % (rank, nsel, nfeat, t_select * 1e3, t_select2 * 1e3))

    def restart():
        L.KLTStopSequentialMode(tc)
        tc.contents.sequentialMode = 1
        capi.arrays_to_featurelist(fl, *sel)

    idx = lambda s: synth.pingpong_index(s, nframes)

    # ---- (1) value: frames resident in HBM, features resident, no host sync ----------
    # (nvidia-smi is sampled every 100 ms from here to the end of the e2e legs; the rows taken are
    # those between the start of the first timed region and the end of the last)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    restart()
    L.KLTB200ResidentBegin(tc, C.c_void_p(d_ptr(0)), 1, ncols, ncols, nrows, fl)
    step = 1
    for _ in range(W):
        L.KLTB200ResidentStep(tc, C.c_void_p(d_ptr(idx(step))), 1, ncols, ncols, nrows)
        step += 1
    live = C.c_ulonglong(0)
    L.klt_dev_live_total(dev, C.byref(live), 1)
    launches0 = L.klt_dev_launch_count(dev)
    barrier()
    wall0 = time.time()
    L.klt_dev_timer_start(dev)
    for _ in range(K):
        L.KLTB200ResidentStep(tc, C.c_void_p(d_ptr(idx(step))), 1, ncols, ncols, nrows)
        step += 1
    ms = C.c_float(0)
    L.klt_dev_timer_stop(dev, C.byref(ms))
    barrier()
    launches = int(L.klt_dev_launch_count(dev) - launches0)
    L.klt_dev_live_total(dev, C.byref(live), 1)
    L.KLTB200ResidentEnd(tc, fl)
    dev_ms, dev_feats = float(ms.value), int(live.value)
    alive_end = int(L.KLTCountRemainingFeatures(fl))

    # ---- (2) e2e: the public KLTTrackFeatures call, pinned HOST frames ---------------
    restart()
    step = 1
    for _ in range(W):
        L.KLTTrackFeatures(tc, C.c_void_p(h_ptr(idx(step - 1))), C.c_void_p(h_ptr(idx(step))), ncols, nrows, fl)
        step += 1
    L.klt_dev_live_total(dev, C.byref(live), 1)          # numerator counted on the device, as in (1)
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        L.KLTTrackFeatures(tc, C.c_void_p(h_ptr(idx(step - 1))), C.c_void_p(h_ptr(idx(step))), ncols, nrows, fl)
        step += 1
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    L.klt_dev_live_total(dev, C.byref(live), 1)
    e2e_feats = int(live.value)

    # ---- (2b) e2e, batched: KLTTrackFeaturesSequence over the same pinned HOST frames -----
    # (one call for K frames, per-frame results into a feature table; frame k+1 crosses PCIe while
    # frame k is processed; the call returns after its single synchronisation)
    restart()
    ft = L.KLTCreateFeatureTable(K + 1, nfeat)
    C.memset(C.cast(ft.contents.feature[0][0], C.c_void_p), 0,          # map the table's pages now
             (K + 1) * nfeat * C.sizeof(capi.KLT_FeatureRec))
    warm = (C.c_void_p * (W + 1))(*[h_ptr(idx(s)) for s in range(W + 1)])
    L.KLTTrackFeaturesSequence(tc, warm, W + 1, ncols, nrows, fl, None, 0, 0)
    seq = (C.c_void_p * (K + 1))(*[h_ptr(idx(W + s)) for s in range(K + 1)])
    L.klt_dev_live_total(dev, C.byref(live), 1)
    barrier()
    t0 = time.perf_counter()
    L.KLTTrackFeaturesSequence(tc, seq, K + 1, ncols, nrows, fl, ft, 0, 0)
    seq_s = time.perf_counter() - t0
    barrier()
    L.klt_dev_live_total(dev, C.byref(live), 1)
    seq_feats = int(live.value)
    L.KLTFreeFeatureTable(ft)
    wall1 = time.time()
    clocks = sampler.stop(wall0, wall1) if sampler else None

    # ---- (3) per-kernel device time (CUDA events on the launching stream) ------------
    restart()
    L.KLTB200ResidentBegin(tc, C.c_void_p(d_ptr(0)), 1, ncols, ncols, nrows, fl)
    step = 1
    for _ in range(min(W, 5)):
        L.KLTB200ResidentStep(tc, C.c_void_p(d_ptr(idx(step))), 1, ncols, ncols, nrows)
        step += 1
    L.klt_dev_profile_begin(dev)
    for _ in range(K):
        L.KLTB200ResidentStep(tc, C.c_void_p(d_ptr(idx(step))), 1, ncols, ncols, nrows)
        step += 1
    L.klt_dev_profile_end(dev)
    prof = L.profile(dev)
    L.KLTB200ResidentEnd(tc, fl)

    # ---- (4) the level-0 kernel alone, launched back to back (no event pair per launch: the
    # brackets of pass (3) add the launch latency that the frame chain hides) ----------------
    l0_b2b_ms = None
    if "l0_fused_kernel" in prof:
        q0 = L.build_desc(tc, ncols, nrows, nlevels_built=1, exact=0)
        nb2b = max(10, min(K, 200))
        for i in range(5):
            L.klt_dev_build(dev, i % 3, C.c_void_p(d_ptr(idx(i))), 1, ncols, C.byref(q0))
        L.klt_dev_timer_start(dev)
        for i in range(nb2b):
            L.klt_dev_build(dev, i % 3, C.c_void_p(d_ptr(idx(i))), 1, ncols, C.byref(q0))
        ms0 = C.c_float(0)
        L.klt_dev_timer_stop(dev, C.byref(ms0))
        l0_b2b_ms = float(ms0.value) / nb2b

    # ---- reduce over ranks: sum of features, max of time -------------------------------
    if world > 1:
        t = torch.tensor([dev_ms, e2e_s, seq_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        c = torch.tensor([dev_feats, e2e_feats, seq_feats], dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        dev_ms, e2e_s, seq_s = float(t[0]), float(t[1]), float(t[2])
        dev_feats, e2e_feats, seq_feats = int(c[0]), int(c[1]), int(c[2])

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        bytes_tab, px = algorithmic_bytes(ncols, nrows, nlevels, ss)
        kernels = {}
        for name, (n, tot) in prof.items():
            per_step = tot / K
            ent = {"launches_per_step": n / K, "ms_per_step": round(per_step, 5)}
            key = name
            if name == "grad_tile" and "l0_fused_kernel" in prof:
                key = "grad_tile_coarse"
            if key in bytes_tab:
                ent["algorithmic_bytes_per_step"] = bytes_tab[key]
                ent["gbs"] = round(bytes_tab[key] / (per_step * 1e-3) / 1e9, 1)
            kernels[name] = ent
        pipe = ["smooth_u8_tile", "grad_tile", "pyrdown_tile", "l0_fused_kernel", "level_fused_kernel[level 1]",
                "level_fused_kernel[level 2]", "level_fused_kernel[level 3+]", "pyramid_mega_kernel"]
        hbm_kernels = {k: v for k, v in kernels.items() if "gbs" in v}
        dom = max(hbm_kernels, key=lambda k: hbm_kernels[k]["ms_per_step"]) if hbm_kernels else None
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if dom and os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(dom)
        roofline = None
        if dom:
            d = kernels[dom]
            roofline = {"kernel": dom, "bound": "hbm", "achieved": d["gbs"], "peak": peak, "unit": "GB/s",
                        "frac": round(d["gbs"] / peak, 4), "traffic": traffic, "peak_source": peak_src,
                        "bytes_per_step": d["algorithmic_bytes_per_step"],
                        "launches_per_step": d["launches_per_step"], "ms_per_step": d["ms_per_step"]}
            if dom == "l0_fused_kernel" and l0_b2b_ms:
                g = d["algorithmic_bytes_per_step"] / (l0_b2b_ms * 1e-3) / 1e9
                roofline["back_to_back"] = {
                    "ms_per_launch": round(l0_b2b_ms, 5), "gbs": round(g, 1), "frac": round(g / peak, 4),
                    "how": "the same kernel launched back to back on distinct frames between ONE pair of "
                           "events (level-0-only builds): its steady-state rate without the per-launch "
                           "event brackets of `kernels`"}
            pipe_ms = sum(v["ms_per_step"] for k, v in kernels.items() if k in pipe)
            if pipe_ms > 0:
                roofline["frame_pipeline"] = {
                    "algorithmic_bytes_per_step": bytes_tab["frame_pipeline"], "ms_per_step": round(pipe_ms, 5),
                    "gbs": round(bytes_tab["frame_pipeline"] / (pipe_ms * 1e-3) / 1e9, 1),
                    "frac": round(bytes_tab["frame_pipeline"] / (pipe_ms * 1e-3) / 1e9 / peak, 4)}
        value = dev_feats / (dev_ms * 1e-3)
        out = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": round(dev_ms / K, 5), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "frames_per_s": round(K * world / (dev_ms * 1e-3), 1),
            "config": {"workload": "synthetic %dx%d translated sequence (2.3,-1.4 px/frame), %d features, "
                                   "%d pyramid levels (subsampling %d), window %dx%d, one independent "
                                   "sequence per GPU" % (ncols, nrows, nfeat, nlevels, ss, window, window),
                       "features_selected": nsel, "features_alive_at_end": alive_end,
                       "l2": "inputs larger than L2: %d distinct frames (%.0f MB) cycled ping-pong; every step "
                             "also writes %.0f MB of new pyramids" % (nframes, nframes * fbytes / 1e6,
                                                                     12 * sum(px) / 1e6),
                       "arithmetic": "fma (default mode); exact mode is the parity-test mode",
                       "select_ms": round(t_select2 * 1e3, 2)},
            "e2e": {"value": round(e2e_feats / e2e_s, 1), "unit": UNIT,
                    "frames_per_s": round(K * world / e2e_s, 1), "ms_per_step": round(e2e_s / K * 1e3, 4),
                    "h2d_bytes_per_step": fbytes + 64 * nfeat, "d2h_bytes_per_step": 12 * nfeat,
                    "api": "KLTTrackFeatures(tc, img1, img2, ncols, nrows, fl) with pinned host frames; the frame "
                           "goes up in 2 bands behind which the pyramid kernels run, the feature records "
                           "(64 B each) are mirrored in one copy and the tracker writes x|y|val (12 B) of "
                           "every feature straight into the caller's pinned feature list"},
            "e2e_sequence": {"value": round(seq_feats / seq_s, 1), "unit": UNIT,
                             "frames_per_s": round(K * world / seq_s, 1), "ms_per_step": round(seq_s / K * 1e3, 4),
                             "h2d_bytes_per_step": fbytes, "d2h_bytes_per_step": 12 * nfeat,
                             "api": "KLTTrackFeaturesSequence(tc, frames[K+1], ..., fl, ft, 0, 0): one call for K "
                                    "pinned host frames, per-frame x|y|val snapshots into a KLT_FeatureTable; the "
                                    "upload of frame k+1 overlaps the kernels of frame k (PCIe-bound), one "
                                    "synchronisation at the end; wall clock around the call"},
            "gpu_launches": launches, "kernels": kernels, "roofline": roofline, "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(capi, fh, ncols, nrows, nfeat, nlevels, ss, window,
                                               budget_s=args.cpu_budget)
        emit_json(out)
    barrier()
    L.KLTFreeFeatureList(fl)
    L.KLTFreeTrackingContext(tc)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- CPU legs
def _ref_library(capi):
    from oracle import oracle_py
    if os.path.exists(oracle_py.REF_PATH):
        return capi.KLTLibrary(oracle_py.REF_PATH), "reference"
    return None, "port"


def _cpu_sequence(capi, frames, ncols, nrows, nfeat, nlevels, ss, window, warmup, steps, budget_s,
                  start_evt=None):
    """select + track on the CPU reference; returns (features, seconds in KLTTrackFeatures, steps)."""
    lib, kind = _ref_library(capi)
    nframes = len(frames)
    idx = lambda s: s % (2 * (nframes - 1)) if s % (2 * (nframes - 1)) < nframes else 2 * (nframes - 1) - s % (2 * (nframes - 1))
    if lib is not None:
        lib.KLTSetVerbosity(0)
        tc = setup_tc(lib, nlevels, ss, window)
        fl = lib.KLTCreateFeatureList(nfeat)
        ptr = lambda i: C.c_void_p(frames[i].ctypes.data)
        lib.KLTSelectGoodFeatures(tc, ptr(0), ncols, nrows, fl)
        track = lambda a, b: lib.KLTTrackFeatures(tc, ptr(a), ptr(b), ncols, nrows, fl)
        count = lambda: lib.KLTCountRemainingFeatures(fl)
    else:                                           # plain-C port (oracle/klt_oracle.c)
        from oracle import oracle_py
        o = oracle_py.Oracle()
        p = o.default_params()
        p.window_width = p.window_height = window
        p.nPyramidLevels, p.subsampling = nlevels, ss
        o.update_border(p)
        state = {"xyv": o.select(frames[0], p, nfeat), "prev": o.build_pyramids(frames[0], p)}

        def track(a, b):
            cur = o.build_pyramids(frames[b], p)
            state["xyv"] = o.track(state["prev"], cur, p, *state["xyv"])
            state["prev"] = cur
        count = lambda: int((state["xyv"][2] >= 0).sum())
    step = 1
    for _ in range(warmup):
        track(idx(step - 1), idx(step))
        step += 1
    if start_evt is not None:
        start_evt()
    feats, secs, done = 0, 0.0, 0
    for _ in range(steps):
        feats += count()
        t0 = time.perf_counter()
        track(idx(step - 1), idx(step))
        secs += time.perf_counter() - t0
        step += 1
        done += 1
        if secs > budget_s:
            break
    return feats, secs, done, kind


def cpu_baseline(capi, frames, ncols, nrows, nfeat, nlevels, ss, window, budget_s=20.0):
    feats, secs, done, kind = _cpu_sequence(capi, frames[:6], ncols, nrows, nfeat, nlevels, ss, window,
                                            warmup=1, steps=1000, budget_s=budget_s)
    return {"value": round(feats / secs, 1), "unit": UNIT, "cores": 1, "kind": kind,
            "frames_per_s": round(done / secs, 3),
            "sample": "1 sequence on 1 host core: KLTSelectGoodFeatures + 1 warm-up + %d timed "
                      "KLTTrackFeatures calls on the same %dx%d frames / %d features (wall clock around "
                      "KLTTrackFeatures only, as the reference driver times it)" % (done, ncols, nrows, nfeat)}


def _ref_worker(args):
    (widx, workload, warmup, steps, budget_s, seed) = args
    pkg = importlib.import_module(PKG)
    synth = importlib.import_module(PKG + ".synth")
    ncols, nrows, nfeat, nlevels, ss, window, _ = WORKLOADS[workload]
    frames = np.empty((6, nrows, ncols), np.uint8)
    for t in range(6):
        synth.frame(ncols, nrows, seed=seed + widx, t=float(t), out=frames[t], threads=1)
    t_begin = [0.0]
    feats, secs, done, kind = _cpu_sequence(pkg.capi, frames, ncols, nrows, nfeat, nlevels, ss, window,
                                            warmup, steps, budget_s,
                                            start_evt=lambda: t_begin.__setitem__(0, time.time()))
    return feats, secs, done, kind, t_begin[0], time.time()


def run_reference(args, rank, world):
    """The reference's own CPU implementation (oracle/_ref = its unmodified sources compiled in
    place), on all host cores: the library is single-threaded, so one independent sequence per core."""
    if rank != 0:
        return
    import multiprocessing as mp
    ncols, nrows, nfeat, nlevels, ss, window, _ = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    per_proc_gb = 0.75 * (ncols * nrows) / (3840 * 2160)
    try:
        avail_gb = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 2 ** 30
    except Exception:
        avail_gb = 16.0
    procs = max(1, min(cores, 64, int(avail_gb * 0.6 / max(per_proc_gb, 1e-3))))
    # a reference step at 4K costs about a second: bound the sample so the run ends in minutes
    budget = args.ref_budget
    warm = min(args.warmup, 1)
    jobs = [(w, args.workload, warm, args.steps, budget, 12345) for w in range(procs)]
    t0 = time.time()
    with mp.get_context("fork").Pool(procs) as pool:
        res = pool.map(_ref_worker, jobs)
    feats = sum(r[0] for r in res)
    frames_done = sum(r[2] for r in res)
    span = max(r[5] for r in res) - min(r[4] for r in res)     # wall time of the timed phase
    busy = max(r[1] for r in res)                               # slowest worker's time inside the calls
    secs = max(busy, 1e-9)
    kind = res[0][3]
    steps_done = min(r[2] for r in res)
    value = feats / secs
    sample = ("%d independent sequences in parallel (one per host core, the library is single-threaded), "
              "each: select + %d warm-up + %d timed KLTTrackFeatures calls at %dx%d / %d features; "
              "timed-phase wall %.1fs" % (procs, warm, steps_done, ncols, nrows, nfeat, span))
    out = {
        "impl": "reference", "metric": METRIC, "value": round(value, 1), "unit": UNIT,
        "n_gpus": args.gpus, "steps": steps_done, "warmup": warm,
        "ms_per_step": round(secs / max(steps_done, 1) * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "frames_per_s": round(frames_done / secs, 3),
        "config": {"workload": "synthetic %dx%d translated sequence (2.3,-1.4 px/frame), %d features, "
                               "%d pyramid levels (subsampling %d), window %dx%d"
                               % (ncols, nrows, nfeat, nlevels, ss, window, window),
                   "requested_steps": args.steps, "requested_warmup": args.warmup},
        "cpu_baseline": {"value": round(value, 1), "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
        "e2e": {"value": round(value, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "total_wall_s": round(time.time() - t0, 1),
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="4k", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--ref-budget", type=float, default=90.0,
                    help="--impl reference: stop a worker's timed loop after this many seconds")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3 and args.impl == "b200":
        log("warmup raised to 3 (timing rules)")
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
