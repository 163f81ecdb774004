"""The batched / hashed / fixed-point minimum-distance walk of enforce_mindist_kernel (restated in
tests/walk_model.py) accepts exactly the candidates of the reference's sequential greedy pass
(src/V1/selectGoodFeatures.c:135-239) -- on random candidate sets, dense clusters, chains, every
mindist class (d < 0, d = 0, d >= 1), pre-stamped maps and short feature lists.  CPU only."""
import numpy as np
import pytest

from tests.walk_model import walk_batched, walk_sequential


def _candidates(rng, w, h, n, kind):
    if kind == "uniform":
        p = rng.permutation(w * h)[:n]
        xs, ys = p % w, p // w
    elif kind == "clusters":                       # blobs of high values, as corners give
        k = max(1, n // 40)
        cx, cy = rng.integers(0, w, k), rng.integers(0, h, k)
        xs = np.clip(cx[rng.integers(0, k, 4 * n)] + rng.integers(-6, 7, 4 * n), 0, w - 1)
        ys = np.clip(cy[rng.integers(0, k, 4 * n)] + rng.integers(-6, 7, 4 * n), 0, h - 1)
        _, first = np.unique(ys * w + xs, return_index=True)      # candidates are distinct pixels
        first = np.sort(first)[:n]
        xs, ys = xs[first], ys[first]
    else:                                          # "chain": a diagonal line, every link within d of the next
        t = np.arange(n)
        xs, ys = (3 * t) % w, (3 * t // w * 3 + (3 * t) % 7) % h
        _, first = np.unique(ys * w + xs, return_index=True)
        first = np.sort(first)
        xs, ys = xs[first], ys[first]
    vals = np.sort(rng.integers(0, 5000, len(xs)))[::-1].copy()   # ranked: descending, ties allowed
    return xs.astype(np.int64), ys.astype(np.int64), vals.astype(np.int64)


@pytest.mark.parametrize("kind", ["uniform", "clusters", "chain"])
@pytest.mark.parametrize("mindist", [0, 1, 2, 5, 10, 40])
def test_batched_walk_equals_sequential_pass(kind, mindist):
    rng = np.random.default_rng(100 * mindist + len(kind))
    w, h = 200, 150
    xs, ys, vals = _candidates(rng, w, h, 9000, kind)
    d = mindist - 1                                 # the reference works with mindist-1 (:157)
    for nopen, min_eig in ((10 ** 9, 1), (37, 1), (500, 2500)):
        ref = walk_sequential(xs, ys, vals, w, h, d, min_eig, nopen)
        rounds = []
        got = walk_batched(xs, ys, vals, w, h, d, min_eig, nopen, stats=rounds)
        assert got == ref
        assert not rounds or max(rounds) <= 32


def test_batched_walk_on_a_prestamped_map():
    """KLTReplaceLostFeatures: the surviving features are stamped before the walk"""
    rng = np.random.default_rng(7)
    w, h, d = 320, 240, 9
    xs, ys, vals = _candidates(rng, w, h, 20000, "clusters")
    fmap = np.zeros((h, w), np.uint8)
    for _ in range(150):
        x, y = int(rng.integers(0, w)), int(rng.integers(0, h))
        fmap[max(0, y - d):y + d + 1, max(0, x - d):x + d + 1] = 1
    ref = walk_sequential(xs, ys, vals, w, h, d, 1, 60, fmap)
    assert walk_batched(xs, ys, vals, w, h, d, 1, 60, fmap) == ref
    assert len(ref) > 0
