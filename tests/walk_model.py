"""Executable statement of the algorithm inside enforce_mindist_kernel (csrc/klt_dev.cu) -- test
infrastructure, numpy / pure Python, no GPU.

The kernel replaces the reference's sequential greedy pass (_enforceMinimumDistance,
src/V1/selectGoodFeatures.c:135-239: walk the candidates in rank order, accept one when its pixel
is not covered, stamp the (2d+1)^2 square around it) by

  * batches of 1024 or 4096 consecutive ranks tested against the featuremap in parallel
    (size of the batch after the next: 1024 while more than a quarter of a batch survives the map),
  * survivors resolved 32 at a time against the candidates accepted earlier in the same batch
    (cell hash: at most one accepted candidate per (d+1)^2 cell, 3 x 3 cells probed),
  * conflicts inside a group of 32 settled by a fixed-point iteration (a lane is accepted once no
    earlier conflicting lane is accepted or undecided, rejected once one is accepted),
  * the list cut at the number of open slots.

`walk_batched` follows those steps literally; tests/test_walk_model.py checks it against
`walk_sequential` on random and adversarial candidate sets, which is the equivalence the kernel's
header comment claims ("identical to the sequential pass for any batch size").
"""
import numpy as np

GB, GK = 1024, 4


def walk_sequential(xs, ys, vals, w, h, d, min_eig, nopen, fmap=None):
    """the reference's pass: returns the accepted candidate indices in order"""
    fmap = np.zeros((h, w), np.uint8) if fmap is None else fmap.copy()
    out = []
    for i in range(len(xs)):
        if len(out) >= nopen:
            break
        if vals[i] < min_eig:
            break                                   # sorted descending: nothing later can pass
        x, y = int(xs[i]), int(ys[i])
        if fmap[y, x]:
            continue
        out.append(i)
        if d >= 0:
            fmap[max(0, y - d):y + d + 1, max(0, x - d):x + d + 1] = 1
    return out


def _settle_group(conf, cand):
    """fixed point on bit masks: conf[l] = earlier lanes of the group within d of lane l"""
    acc, und = 0, cand
    rounds = 0
    while und:
        okm = rejm = 0
        for l in range(32):
            if not (und >> l) & 1:
                continue
            c = conf[l] & cand
            if c & acc:
                rejm |= 1 << l
            elif not (c & und):
                okm |= 1 << l
        assert okm | rejm, "no progress"
        acc |= okm
        und &= ~(okm | rejm)
        rounds += 1
    return acc, rounds


def walk_batched(xs, ys, vals, w, h, d, min_eig, nopen, fmap=None, stats=None):
    fmap = np.zeros((h, w), np.uint8) if fmap is None else fmap.copy()
    n = len(xs)
    out = []
    done = nopen == 0
    spaced = d >= 1
    cs = d + 1 if spaced else 1
    gk = gk_next = gk_want = 1                      # (the kernel starts replacement walks at GK)
    base = 0
    while base < n and not done:
        gk_next = gk_want                           # the next batch's keys are fetched now, in this layout
        bsz = GB * gk
        idx = np.arange(base, min(base + bsz, n))
        alive = (vals[idx] >= min_eig) & (fmap[ys[idx], xs[idx]] == 0)
        surv = idx[alive]                           # rank order
        if vals[base] < min_eig:
            done = True
        table = {}                                  # cell -> accepted candidate of this batch
        budget = nopen - len(out)
        accepted = []
        for g0 in range(0, len(surv), 32):
            if len(accepted) >= budget:
                break
            grp = surv[g0:g0 + 32]
            hit = [False] * 32
            conf = [0] * 32
            for l, i in enumerate(grp):
                px, py = int(xs[i]), int(ys[i])
                if spaced:
                    cx, cy = px // cs, py // cs
                    for dy in (-1, 0, 1):
                        for dx in (-1, 0, 1):
                            a = table.get((cx + dx, cy + dy))
                            if a is not None and abs(px - a[0]) <= d and abs(py - a[1]) <= d:
                                hit[l] = True
                    for m in range(l):
                        j = grp[m]
                        if abs(px - int(xs[j])) <= d and abs(py - int(ys[j])) <= d:
                            conf[l] |= 1 << m
            cand = sum(1 << l for l in range(len(grp)) if not hit[l])
            acc, rounds = _settle_group(conf, cand)
            if stats is not None:
                stats.append(rounds)
            while bin(acc).count("1") > budget - len(accepted):
                acc &= ~(1 << (acc.bit_length() - 1))           # list full: the first ones in rank order
            for l, i in enumerate(grp):
                if (acc >> l) & 1:
                    accepted.append(i)
                    if spaced:
                        key = (int(xs[i]) // cs, int(ys[i]) // cs)
                        assert key not in table, "two accepted candidates in one cell"
                        table[key] = (int(xs[i]), int(ys[i]))
        for i in accepted:
            out.append(i)
            x, y = int(xs[i]), int(ys[i])
            if d >= 0:
                fmap[max(0, y - d):y + d + 1, max(0, x - d):x + d + 1] = 1
        if len(out) >= nopen:
            done = True
        gk_want = 1 if len(surv) * 4 > bsz else GK
        base += bsz
        gk = gk_next
    return out
