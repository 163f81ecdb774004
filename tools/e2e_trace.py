#!/usr/bin/env python
"""Device timeline of the synchronous KLTTrackFeatures call (bench.py's e2e leg): every copy and
kernel bracketed by CUDA events on its own stream (klt_dev_profile_* / klt_dev_trace_get).
  python tools/e2e_trace.py [--steps 3] [--workload 4k] [--pageable] [--register]
--pageable: the frames come from two ordinary host buffers refilled before every call (what the
reference's driver passes); --register: ... and the library may page-lock them in place.
KLT_B200_TIMING=1 in the environment adds the library's own host-phase averages (per 64 calls).
The event brackets perturb the run a little; use it to see what overlaps what, not for the totals."""
import argparse, ctypes as C, importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--workload", default="4k")
    ap.add_argument("--pageable", action="store_true")
    ap.add_argument("--register", action="store_true")
    ap.add_argument("--calls", type=int, default=128)
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
    L.klt_dev_set_register_frames(dev, 1 if a.register else 0)
    frames = torch.empty((nframes, nrows, ncols), dtype=torch.uint8, pin_memory=True)
    fh = frames.numpy()
    bench.make_frames(synth, ncols, nrows, nframes, 12345, fh)
    pinned = lambda i: C.c_void_p(frames.data_ptr() + i * ncols * nrows)
    bufs = [np.empty((nrows, ncols), np.uint8) for _ in range(2)]
    fl = L.KLTCreateFeatureList(nfeat)
    L.KLTSelectGoodFeatures(tc, pinned(0), ncols, nrows, fl)
    idx = lambda s: synth.pingpong_index(s, nframes)
    np.copyto(bufs[0], fh[0])

    def call(step):
        if a.pageable:
            cur, prv = bufs[step & 1], bufs[(step - 1) & 1]
            np.copyto(cur, fh[idx(step)])
            p1, p2 = C.c_void_p(prv.ctypes.data), C.c_void_p(cur.ctypes.data)
        else:
            p1, p2 = pinned(idx(step - 1)), pinned(idx(step))
        t0 = time.perf_counter()
        L.KLTTrackFeatures(tc, p1, p2, ncols, nrows, fl)
        return time.perf_counter() - t0

    step = 1
    t_warm = time.perf_counter()
    while time.perf_counter() - t_warm < 1.0:          # clocks and link at speed before anything is timed
        call(step); step += 1
    tot = 0.0
    for _ in range(a.calls):
        tot += call(step); step += 1
    print("un-instrumented: %.1f us / call (%s frames%s), %d host buffers page-locked in place"
          % (tot / a.calls * 1e6, "pageable" if a.pageable else "pinned", ", register" if a.register else "",
             L.klt_dev_registered_host_frames(dev)))
    L.klt_dev_profile_begin(dev)
    host = []
    for _ in range(a.steps):
        host.append(call(step) * 1e6); step += 1
    L.klt_dev_profile_end(dev)
    tr = L.trace(dev)
    print("host us per instrumented call:", ["%.1f" % h for h in host])
    # the last call: everything after the last copy_d2h-terminated group
    ends = [i for i, r in enumerate(tr) if r[0] == "track7w_kernel"]
    last = tr[ends[-2] + 1:] if len(ends) >= 2 else tr
    o = min(r[1] for r in last)
    for name, s, e in sorted(last, key=lambda r: r[1]):
        print("%-28s %8.1f -> %8.1f us  (%6.1f)" % (name, (s - o) * 1e3, (e - o) * 1e3, (e - s) * 1e3))
    L.KLTFreeFeatureList(fl); L.KLTFreeTrackingContext(tc)


if __name__ == "__main__":
    main()
