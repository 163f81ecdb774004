#!/usr/bin/env python
"""Builds of distinct resident 4K frames, back to back (the launches ncu profiles).
   python tools/l0_only.py [launches] [levels]"""
import ctypes as C, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
pkg = importlib.import_module(bench.PKG)
rt = importlib.import_module(bench.PKG + ".runtime")
synth = importlib.import_module(bench.PKG + ".synth")
L = rt.load(); L.require_gpu(); L.KLTSetVerbosity(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
levels = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ncols, nrows, nfeat, nlevels, ss, window, _ = bench.WORKLOADS["4k"]
tc = bench.setup_tc(L, nlevels, ss, window, device=0)
dev = L.KLTB200Device(tc)
frames = torch.empty((8, nrows, ncols), dtype=torch.uint8)
bench.make_frames(synth, ncols, nrows, 8, 12345, frames.numpy())
fd = frames.cuda()
q = L.build_desc(tc, ncols, nrows, nlevels_built=levels, exact=0)
for i in range(4):
    L.klt_dev_build(dev, i % 3, C.c_void_p(fd.data_ptr() + (i % 8) * ncols * nrows), 1, ncols, C.byref(q))
L.klt_dev_timer_start(dev)
for i in range(n):
    L.klt_dev_build(dev, i % 3, C.c_void_p(fd.data_ptr() + (i % 8) * ncols * nrows), 1, ncols, C.byref(q))
ms = C.c_float(0)
L.klt_dev_timer_stop(dev, C.byref(ms))
print("%d builds of %d level(s): %.2f us each" % (n, levels, ms.value / n * 1e3))
