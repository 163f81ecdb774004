#!/usr/bin/env python
"""Device time of klt_dev_build alone (resident 4K frame): mega kernel vs per-level kernels,
level 0 only vs full pyramid."""
import ctypes as C, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa
import torch
pkg = importlib.import_module(bench.PKG)
rt = importlib.import_module(bench.PKG + ".runtime")
synth = importlib.import_module(bench.PKG + ".synth")
L = rt.load(); L.require_gpu(); L.KLTSetVerbosity(0)
ncols, nrows, nfeat, nlevels, ss, window, _ = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "4k"]
tc = bench.setup_tc(L, nlevels, ss, window, device=0)
dev = L.KLTB200Device(tc)
nf = 12
frames = torch.empty((nf, nrows, ncols), dtype=torch.uint8, pin_memory=True)
bench.make_frames(synth, ncols, nrows, nf, 1, frames.numpy())
fd = frames.cuda()
for nb in (1, nlevels):
    q = L.build_desc(tc, ncols, nrows, nlevels_built=nb, exact=0)
    for mega in (0, 1):
        L.klt_dev_disable_mega(dev, 1 - mega)
        for i in range(5):
            L.klt_dev_build(dev, i % 3, C.c_void_p(fd.data_ptr() + (i % nf) * ncols * nrows), 1, ncols, C.byref(q))
        L.klt_dev_timer_start(dev)
        N = 60
        for i in range(N):
            L.klt_dev_build(dev, i % 3, C.c_void_p(fd.data_ptr() + (i % nf) * ncols * nrows), 1, ncols, C.byref(q))
        ms = C.c_float(0); L.klt_dev_timer_stop(dev, C.byref(ms))
        print("levels %d mega %d: %.1f us / build (mega flag %d)" % (nb, mega, ms.value / N * 1e3, L.klt_dev_last_build_mega(dev)))
