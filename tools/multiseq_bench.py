#!/usr/bin/env python
"""tools/multiseq_bench.py -- BASELINE config 5: 64 independent synthetic 1920x1080 sequences,
1024 features each, partitioned over the GPUs of one box.

  python tools/multiseq_bench.py [--sequences 64] [--frames 17] [--threads 4]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P tools/multiseq_bench.py

Sequence s runs on rank s mod G with its own tracking context (own streams and pyramids); there is
no collective on the data path.  Every sequence is one KLTTrackFeaturesSequence call (include/
klt_b200.h) over pinned host frames into its own KLT_FeatureTable; a rank drives `--threads`
sequences at a time from host threads so that one sequence's PCIe transfer overlaps another's
kernels.  The tables are gathered on rank 0 through torch.distributed (gather_object) and the
counters reduced as (sum of features, max of seconds).  Selection (frame 0 of every sequence) is
done before the timed region, as the reference driver times only KLTTrackFeatures.

One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "klt-feature-tracker-acceleration-gpus_b200"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sequences", type=int, default=64)
    ap.add_argument("--frames", type=int, default=17, help="frames per sequence (16 tracked)")
    ap.add_argument("--features", type=int, default=1024)
    ap.add_argument("--threads", type=int, default=4, help="sequences in flight per GPU")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import torch
    import torch.distributed as dist
    pkg = importlib.import_module(PKG)
    rt = importlib.import_module(PKG + ".runtime")
    synth = importlib.import_module(PKG + ".synth")
    multiseq = importlib.import_module(PKG + ".multiseq")
    capi = pkg.capi
    L = rt.load()
    L.require_gpu()
    L.KLTSetVerbosity(0)
    torch.cuda.set_device(local_rank)
    json_fd = 1
    if world > 1:
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)                       # NCCL banners go to stderr, the JSON line to the real stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    W, H, NF, T = 1920, 1080, args.features, args.frames
    mine = multiseq.my_sequences(args.sequences, rank, world)
    rng = np.random.default_rng(2024)
    vel = rng.uniform(-3.0, 3.0, size=(args.sequences, 2))      # per-sequence velocity, px / frame

    # ---- data: pinned host frames of this rank's sequences --------------------------------
    t0 = time.time()
    frames = torch.empty((len(mine), T, H, W), dtype=torch.uint8, pin_memory=True)
    fh = frames.numpy()
    for k, s in enumerate(mine):
        for t in range(T):
            synth.frame(W, H, seed=1000 + s, t=float(t), velocity=tuple(vel[s]), out=fh[k, t])
    print("[rank %d] %d sequences x %d frames generated in %.1fs" % (rank, len(mine), T, time.time() - t0),
          file=sys.stderr)

    # ---- contexts, selection (untimed) ----------------------------------------------------------
    ctx = []
    for k, s in enumerate(mine):
        tc = L.KLTCreateTrackingContext()          # tc defaults: 2 levels, subsampling 4, 7x7 window
        tc.contents.sequentialMode = 1
        L.KLTB200SetDevice(tc, local_rank)
        fl = L.KLTCreateFeatureList(NF)
        ft = L.KLTCreateFeatureTable(T, NF)
        C.memset(C.cast(ft.contents.feature[0][0], C.c_void_p), 0, T * NF * C.sizeof(capi.KLT_FeatureRec))
        L.KLTSelectGoodFeatures(tc, C.c_void_p(frames[k, 0].data_ptr()), W, H, fl)
        L.KLTStoreFeatureList(fl, ft, 0)
        ptrs = (C.c_void_p * T)(*[frames[k, t].data_ptr() for t in range(T)])
        ctx.append((tc, fl, ft, ptrs))

    def run_one(k):
        tc, fl, ft, ptrs = ctx[k]
        L.KLTTrackFeaturesSequence(tc, ptrs, T, W, H, fl, ft, 0, 0)
        return k

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up: the first two frames of every sequence's first slot (module load, allocations)
    if ctx:
        tc, fl, ft, ptrs = ctx[0]
        keep = capi.featurelist_to_arrays(fl)
        L.KLTTrackFeaturesSequence(tc, ptrs, 3, W, H, fl, None, 0, 0)
        L.KLTStopSequentialMode(tc)
        tc.contents.sequentialMode = 1
        capi.arrays_to_featurelist(fl, *keep)

    barrier()
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=max(1, args.threads)) as pool:
        list(pool.map(run_one, range(len(ctx))))
    torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    barrier()

    # ---- results: tables -> numpy, host gather, counters ---------------------------------------
    local, feats = {}, 0
    words = C.sizeof(capi.KLT_FeatureRec) // 4
    for k, s in enumerate(mine):
        ft = ctx[k][2]                      # the table's records are one [feature][frame] block
        base = C.cast(ft.contents.feature[0][0], C.c_void_p).value
        ri = np.ctypeslib.as_array(C.cast(base, C.POINTER(C.c_int32)), shape=(NF, T, words))
        rf = np.ctypeslib.as_array(C.cast(base, C.POINTER(C.c_float)), shape=(NF, T, words))
        val = ri[:, :, 2]
        feats += int((val[:, :-1] >= 0).sum())      # features entering each tracked frame
        local[s] = np.stack([rf[:, :, 0], rf[:, :, 1], val.astype(np.float32)], axis=-1)
    tables = multiseq.gather_tables(local, rank, world, dist if world > 1 else None)
    tot_feats, max_secs = multiseq.aggregate(feats, secs, world, dist if world > 1 else None,
                                             device="cuda" if world > 1 else None)
    if rank == 0:
        assert len(tables) == args.sequences
        alive = [int((t[:, -1, 2] >= 0).sum()) for t in tables.values()]
        out = {"metric": "tracked_features_per_s", "value": round(tot_feats / max_secs, 1), "unit": "features/s",
               "frames_per_s": round(args.sequences * (T - 1) / max_secs, 1), "n_gpus": world,
               "seconds": round(max_secs, 4), "scaling": "strong",
               "config": {"workload": "BASELINE config 5: %d independent synthetic %dx%d sequences, %d features "
                                      "each, %d tracked frames each, tc defaults (2 levels, subsampling 4, 7x7), "
                                      "sequence s on GPU s mod G, %d sequences in flight per GPU"
                                      % (args.sequences, W, H, NF, T - 1, args.threads),
                          "api": "KLTTrackFeaturesSequence per sequence, pinned host frames, feature tables "
                                 "gathered on rank 0 (host)"},
               "features_alive_at_end_min_max": [min(alive), max(alive)], "data": "synthetic"}
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    for tc, fl, ft, _ in ctx:
        L.KLTFreeFeatureTable(ft)
        L.KLTFreeFeatureList(fl)
        L.KLTFreeTrackingContext(tc)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
