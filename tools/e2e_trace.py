#!/usr/bin/env python
"""Device timeline of the synchronous KLTTrackFeatures call (bench.py's e2e leg): every copy and
kernel bracketed by CUDA events on its own stream (klt_dev_profile_* / klt_dev_trace_get).
  python tools/e2e_trace.py [--steps 3] [--workload 4k]
The event brackets perturb the run a little; use it to see what overlaps what, not for the totals."""
import argparse, ctypes as C, importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--workload", default="4k")
    a = ap.parse_args()
    import torch
    pkg = importlib.import_module(bench.PKG)
    rt = importlib.import_module(bench.PKG + ".runtime")
    synth = importlib.import_module(bench.PKG + ".synth")
    L = rt.load(); L.require_gpu(); L.KLTSetVerbosity(0)
    ncols, nrows, nfeat, nlevels, ss, window, _ = bench.WORKLOADS[a.workload]
    nframes = 6
    tc = bench.setup_tc(L, nlevels, ss, window, device=0)
    dev = L.KLTB200Device(tc)
    frames = torch.empty((nframes, nrows, ncols), dtype=torch.uint8, pin_memory=True)
    bench.make_frames(synth, ncols, nrows, nframes, 12345, frames.numpy())
    ptr = lambda i: C.c_void_p(frames.data_ptr() + i * ncols * nrows)
    fl = L.KLTCreateFeatureList(nfeat)
    L.KLTSelectGoodFeatures(tc, ptr(0), ncols, nrows, fl)
    idx = lambda s: synth.pingpong_index(s, nframes)
    step = 1
    for _ in range(10):
        L.KLTTrackFeatures(tc, ptr(idx(step - 1)), ptr(idx(step)), ncols, nrows, fl); step += 1
    t0 = time.perf_counter()
    for _ in range(20):
        L.KLTTrackFeatures(tc, ptr(idx(step - 1)), ptr(idx(step)), ncols, nrows, fl); step += 1
    print("un-instrumented: %.1f us / call" % ((time.perf_counter() - t0) / 20 * 1e6))
    L.klt_dev_profile_begin(dev)
    host = []
    for _ in range(a.steps):
        h0 = time.perf_counter()
        L.KLTTrackFeatures(tc, ptr(idx(step - 1)), ptr(idx(step)), ncols, nrows, fl); step += 1
        host.append((time.perf_counter() - h0) * 1e6)
    L.klt_dev_profile_end(dev)
    tr = L.trace(dev)
    print("host us per instrumented call:", ["%.1f" % h for h in host])
    # split per call: each call starts with the feature copy_h2d
    calls, cur = [], []
    for rec in tr:
        if rec[0] == "copy_h2d" and cur and cur[0][0] == "copy_h2d" and len(cur) > 1 and any(r[0] == "copy_d2h" for r in cur):
            calls.append(cur); cur = []
        cur.append(rec)
    calls.append(cur)
    for c in calls[-1:]:
        o = min(r[1] for r in c)
        for name, s, e in sorted(c, key=lambda r: r[1]):
            print("%-22s %8.1f -> %8.1f us  (%6.1f)" % (name, (s - o) * 1e3, (e - o) * 1e3, (e - s) * 1e3))
    L.KLTFreeFeatureList(fl); L.KLTFreeTrackingContext(tc)


if __name__ == "__main__":
    main()
