#!/usr/bin/env python
"""BASELINE configs 2 and 3 across their FULL sequences on the GPU, against the oracle.

  config 2: data/images_traffic, img1..img551.pgm, 1000 features, no replacement (550 steps)
  config 3: data/images_laptops, img1..img1003.pgm, 2000 features, KLTReplaceLostFeatures after
            every frame (1002 steps)            (reference driver loop src/V3/example3.c:54-76)

For each: teacher-forced parity of every step in both arithmetic modes (exact: bit for bit; fma:
north_star gate, every disagreement explained or the run fails), replacement bit for bit in both
modes, and the free-running drift table (exact mode must be 0 everywhere).  See
tests/longrun_common.py.  The datasets (462 MB) are not part of the repo: point --data at a
directory holding images_traffic/ and images_laptops/ (default: data_full/ in the repo root --
git-ignored -- or /root/reference/data).

  python tools/full_sequence_report.py [--data DIR] [--frames N] [--out profiles/r2_full_sequences.json]
"""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from tests import longrun_common as lr  # noqa: E402
from tests.gpu_common import params_from_tc  # noqa: E402

PKG = "klt-feature-tracker-acceleration-gpus_b200"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data", default=None)
    ap.add_argument("--frames", type=int, default=0, help="limit the number of frames (0: all)")
    ap.add_argument("--only", default="", help="config2 or config3")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "full_sequences.json"))
    args = ap.parse_args()
    data = args.data
    for cand in (os.path.join(ROOT, "data_full"), "/root/reference/data"):
        if data is None and os.path.isdir(os.path.join(cand, "images_traffic")):
            data = cand
    if data is None:
        raise SystemExit("no dataset directory found (see --data)")
    pkg = importlib.import_module(PKG)
    rt = importlib.import_module(PKG + ".runtime")
    from oracle import oracle_py
    capi = pkg.capi
    L = rt.load(); L.require_gpu(); L.KLTSetVerbosity(0)
    oracle = oracle_py.Oracle()
    report = {"data": data, "configs": {}}
    for name, dataset, last, n, replace in (("config2", "images_traffic", 551, 1000, False),
                                            ("config3", "images_laptops", 1003, 2000, True)):
        if args.only and args.only != name:
            continue
        if args.frames:
            last = min(last, args.frames)
        frames = [capi.read_pgm_numpy(os.path.join(data, dataset, "img%d.pgm" % i)) for i in range(1, last + 1)]
        tc = L.KLTCreateTrackingContext()
        p = params_from_tc(oracle, tc)
        L.KLTFreeTrackingContext(tc)
        t0 = time.time()
        table, tracked = lr.oracle_free_run(oracle, oracle_py, frames, p, n, replace)
        t_oracle = time.time() - t0
        lost = np.bincount(-tracked[2][tracked[2] < 0], minlength=6).tolist()
        entry = {"dataset": dataset, "frames": len(frames), "features": n, "replace": replace,
                 "oracle_free_run_s": round(t_oracle, 1),
                 "features_found_on_first_frame": int((table[2][0] > 0).sum()),
                 "alive_at_end_oracle": int((table[2][-1] >= 0).sum()),
                 "tracker_status_cells": {"NOT_FOUND": lost[1], "SMALL_DET": lost[2], "MAX_ITERATIONS": lost[3],
                                          "OOB": lost[4], "LARGE_RESIDUE": lost[5]}}
        for exact in (1, 0):
            mode = "exact" if exact else "fma"
            t0 = time.time()
            entry["teacher_forced_" + mode] = lr.teacher_forced(L, capi, oracle, frames, n, exact, replace, table, tracked)
            entry["teacher_forced_" + mode]["wall_s"] = round(time.time() - t0, 1)
            gpu = lr.gpu_free_run(L, capi, frames, n, exact, replace)
            d = lr.drift_table(table, gpu)
            entry["free_running_" + mode] = d
            if exact:
                assert d["total"]["identical_cells"] == 1.0, d["total"]
            print(name, mode, json.dumps(entry["teacher_forced_" + mode]), json.dumps(d["total"]), flush=True)
        report["configs"][name] = entry
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(report, fh, indent=1)
    print("written", args.out)


if __name__ == "__main__":
    main()
