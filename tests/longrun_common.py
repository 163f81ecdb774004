"""Whole-sequence parity runs (BASELINE configs 2 and 3 "across the full sequence").

Three views of one sequence, all against the oracle (or oracle/_ref) driven by the reference's own
driver loop -- select on frame 0, then per frame KLTTrackFeatures [+ KLTReplaceLostFeatures]
(reference src/V3/example3.c:54-76):

  oracle_free_run   the oracle's own free-running feature table (the teacher)
  teacher_forced    frame k starts from the ORACLE's list of frame k-1; the GPU's one step is
                    compared with the oracle's one step (exact: bit for bit; fma: north_star gate),
                    replacement included -- rounding differences cannot compound, so every
                    disagreement is local and is checked against the thresholds
  gpu_free_run      the GPU left alone for the whole sequence (KLTTrackFeaturesSequence); in exact
                    mode its table must equal the oracle's bit for bit, in fma mode the drift
                    between the two is REPORTED (drift_table), not gated: the survey measured that
                    two CPU builds of the reference itself (with / without FMA contraction) drift
                    apart by up to 2.8 px over 550 frames.

Used by tests/test_gpu_longrun.py (synthetic 640x480, travels with the repo) and by
tools/full_sequence_report.py (the real 551- and 1003-frame datasets, when present).
"""
import numpy as np

from tests.gpu_common import check_fma_step, params_from_tc


def oracle_free_run(oracle, oracle_mod, frames, p, n, replace):
    """-> x, y, val arrays [nframes, n] (row 0 = the selection), and the per-frame lists BEFORE
    replacement (what the tracker returned), needed to teacher-force the replacement step"""
    nf = len(frames)
    X = np.zeros((nf, n), np.float32); Y = np.zeros((nf, n), np.float32); V = np.zeros((nf, n), np.int32)
    TX = np.zeros((nf, n), np.float32); TY = np.zeros((nf, n), np.float32); TV = np.zeros((nf, n), np.int32)
    x, y, v = oracle.select(frames[0], p, n, sort_kind=oracle_mod.SORT_STABLE)
    X[0], Y[0], V[0] = x, y, v
    TX[0], TY[0], TV[0] = x, y, v
    prev = oracle.build_pyramids(frames[0], p)
    for k in range(1, nf):
        cur = oracle.build_pyramids(frames[k], p)
        x, y, v = oracle.track(prev, cur, p, x, y, v)
        TX[k], TY[k], TV[k] = x, y, v
        if replace:
            x, y, v = oracle.select(frames[k], p, n, sort_kind=oracle_mod.SORT_STABLE, replace=True,
                                    last=cur, x=x, y=y, val=v)
        X[k], Y[k], V[k] = x, y, v
        prev = cur
    return (X, Y, V), (TX, TY, TV)


def teacher_forced(L, capi, oracle, frames, n, exact, replace, table, tracked, tc_setup=None, every=1):
    """One GPU step per frame from the oracle's state.  Returns a report dict; raises on an
    unexplained disagreement."""
    X, Y, V = table
    TX, TY, TV = tracked
    tc = L.KLTCreateTrackingContext()
    tc.contents.sequentialMode = 1
    if tc_setup:
        tc_setup(tc)
    L.KLTB200SetExact(tc, exact)
    p = params_from_tc(oracle, tc)
    fl = L.KLTCreateFeatureList(n)
    L.select(tc, frames[0], fl)
    gx, gy, gv = capi.featurelist_to_arrays(fl)
    assert np.array_equal(gv, V[0]) and np.array_equal(gx, X[0]) and np.array_equal(gy, Y[0]), "selection differs"
    rep = {"frames": len(frames) - 1, "steps_checked": 0, "features_entering": 0, "status_disagreements": 0,
           "explained": 0, "max_err_px": 0.0, "replace_steps": 0, "replaced": 0}
    prev = oracle.build_pyramids(frames[0], p) if not exact else None
    for k in range(1, len(frames)):
        capi.arrays_to_featurelist(fl, X[k - 1], Y[k - 1], V[k - 1])        # teacher forcing
        L.track(tc, frames[k - 1], frames[k], fl)
        gx, gy, gv = capi.featurelist_to_arrays(fl)
        live = V[k - 1] >= 0
        rep["features_entering"] += int(live.sum())
        if exact:
            assert np.array_equal(gv, TV[k]), "frame %d: status differs at %s" % (k, np.nonzero(gv != TV[k])[0][:8])
            assert gx.tobytes() == TX[k].tobytes() and gy.tobytes() == TY[k].tobytes(), "frame %d: positions differ" % k
        else:
            cur = oracle.build_pyramids(frames[k], p)
            if k % every == 0 or not np.array_equal(gv, TV[k]):
                rep["explained"] += check_fma_step(oracle, p, prev, cur, X[k - 1], Y[k - 1], V[k - 1], gx, gy, gv,
                                                   TX[k], TY[k], TV[k], "frame %d" % k)
            rep["status_disagreements"] += int((gv != TV[k]).sum())
            both = (gv == TV[k]) & (gv >= 0)
            if both.any():
                rep["max_err_px"] = max(rep["max_err_px"], float(np.maximum(np.abs(gx - TX[k]), np.abs(gy - TY[k]))[both].max()))
            prev = cur
        rep["steps_checked"] += 1
        if replace:
            # the replacement step from the ORACLE's tracked list: integer keys, so it must be
            # identical in BOTH arithmetic modes (level 0 is rebuilt in exact arithmetic for it)
            capi.arrays_to_featurelist(fl, TX[k], TY[k], TV[k])
            L.replace(tc, frames[k], fl)
            rx, ry, rv = capi.featurelist_to_arrays(fl)
            assert np.array_equal(rv, V[k]) and np.array_equal(rx, X[k]) and np.array_equal(ry, Y[k]), \
                "frame %d: replacement differs in %d slots" % (k, int(((rv != V[k]) | (rx != X[k]) | (ry != Y[k])).sum()))
            rep["replace_steps"] += 1
            rep["replaced"] += int(((TV[k] < 0) & (V[k] > 0)).sum())
    L.KLTFreeFeatureList(fl)
    L.KLTFreeTrackingContext(tc)
    return rep


def gpu_free_run(L, capi, frames, n, exact, replace, tc_setup=None):
    """the GPU on its own for the whole sequence, through the batched driver call"""
    tc = L.KLTCreateTrackingContext()
    tc.contents.sequentialMode = 1
    if tc_setup:
        tc_setup(tc)
    L.KLTB200SetExact(tc, exact)
    fl = L.KLTCreateFeatureList(n)
    ft = L.KLTCreateFeatureTable(len(frames), n)
    L.select(tc, frames[0], fl)
    L.KLTStoreFeatureList(fl, ft, 0)
    L.track_sequence(tc, frames, fl, ft, 0, replace)
    tab = capi.featuretable_to_array(ft)               # [n, nframes]
    L.KLTFreeFeatureTable(ft)
    L.KLTFreeFeatureList(fl)
    L.KLTFreeTrackingContext(tc)
    return (np.ascontiguousarray(tab["x"].T), np.ascontiguousarray(tab["y"].T), np.ascontiguousarray(tab["val"].T))


def drift_table(oracle_tab, gpu_tab, checkpoints=None):
    """free-running drift, cell by cell (slot i at frame k): status agreement and the distribution
    of the coordinate differences among cells alive on both sides"""
    OX, OY, OV = oracle_tab
    GX, GY, GV = gpu_tab
    nf = OX.shape[0]
    if checkpoints is None:
        checkpoints = sorted(set([1, 10, 50, 100, 200, 300, 400, 500, 550, 750, 1000, nf - 1]))
    rows = []
    for k in checkpoints:
        if k < 1 or k >= nf:
            continue
        alive_o, alive_g = OV[k] >= 0, GV[k] >= 0
        both = alive_o & alive_g
        d = np.maximum(np.abs(OX[k] - GX[k]), np.abs(OY[k] - GY[k]))[both]
        rows.append({"frame": int(k), "alive_oracle": int(alive_o.sum()), "alive_gpu": int(alive_g.sum()),
                     "alive_both": int(both.sum()),
                     "status_agree": float(((OV[k] >= 0) == (GV[k] >= 0)).mean()),
                     "max_px": float(d.max()) if d.size else 0.0,
                     "median_px": float(np.median(d)) if d.size else 0.0,
                     "frac_gt_0.01px": float((d > 0.01).mean()) if d.size else 0.0,
                     "frac_gt_0.1px": float((d > 0.1).mean()) if d.size else 0.0,
                     "frac_gt_1px": float((d > 1.0).mean()) if d.size else 0.0})
    # With replacement the SLOTS stop corresponding as soon as one run loses a feature one frame
    # earlier than the other (the refill lands elsewhere), although both keep following the same
    # image points: compare the two point SETS as well (nearest oracle point of every GPU point;
    # features are >= mindist apart, so the match is unambiguous).
    try:
        from scipy.spatial import cKDTree
        for row in rows:
            k = row["frame"]
            ao, ag = OV[k] >= 0, GV[k] >= 0
            if ao.any() and ag.any():
                dist, _ = cKDTree(np.stack([OX[k][ao], OY[k][ao]], 1)).query(np.stack([GX[k][ag], GY[k][ag]], 1))
                row["set_match_within_0.01px"] = float((dist <= 0.01).mean())
                row["set_match_within_0.1px"] = float((dist <= 0.1).mean())
                row["set_match_within_1px"] = float((dist <= 1.0).mean())
    except ImportError:
        pass
    alive_o, alive_g = OV[1:] >= 0, GV[1:] >= 0
    both = alive_o & alive_g
    d = np.maximum(np.abs(OX[1:] - GX[1:]), np.abs(OY[1:] - GY[1:]))[both]
    total = {"cells": int(OV[1:].size), "cells_alive_both": int(both.sum()),
             "status_agree": float((alive_o == alive_g).mean()),
             "identical_cells": float(((OX[1:] == GX[1:]) & (OY[1:] == GY[1:]) & (OV[1:] == GV[1:])).mean()),
             "max_px": float(d.max()) if d.size else 0.0,
             "frac_gt_0.01px": float((d > 0.01).mean()) if d.size else 0.0,
             "frac_gt_0.1px": float((d > 0.1).mean()) if d.size else 0.0,
             "frac_gt_1px": float((d > 1.0).mean()) if d.size else 0.0}
    return {"checkpoints": rows, "total": total}
