import subprocess, sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASE = r'''
import sys, importlib, numpy as np
sys.path.insert(0, %r)
from tests.conftest import synth_image
pkg = importlib.import_module("klt-feature-tracker-acceleration-gpus_b200")
L = importlib.import_module(pkg.__name__ + ".runtime").load()
L.KLTSetVerbosity(0)
nlev, ss, h, w, exact, nb, reps = %d, %d, %d, %d, %d, %d, %d
tc = L.KLTCreateTrackingContext()
tc.contents.nPyramidLevels, tc.contents.subsampling = nlev, ss
L.KLTUpdateTCBorder(tc)
dev = L.KLTB200Device(tc)
img = synth_image(w, h, seed=1)
q = L.build_desc(tc, w, h, nlevels_built=nb, exact=exact)
for r in range(reps):
    L.dev_build(dev, r %% 3, img, q)
print("ok mega=%%d" %% L.klt_dev_last_build_mega(dev))
'''
cases = [(2, 4, 240, 320, 0, 2, 1), (2, 4, 240, 320, 1, 2, 1), (2, 4, 240, 320, 1, 2, 5),
         (3, 2, 240, 320, 1, 3, 3), (3, 2, 240, 320, 0, 3, 3), (2, 4, 480, 640, 0, 2, 3), (4, 2, 1080, 1920, 0, 4, 6)]
for envs in ({}, {"KLT_B200_MEGA_SERIAL": "1"}, {"KLT_B200_MEGA_COARSE_EVERY": "100000"}):
    for c in cases:
        t0 = time.time()
        p = subprocess.run([sys.executable, "-c", CASE % ((ROOT,) + c)], capture_output=True, text=True, timeout=120,
                           env=dict(os.environ, **envs))
        print(envs, c, "%.1fs" % (time.time() - t0), p.stdout.strip()[-300:], p.stderr.strip()[-60:])
