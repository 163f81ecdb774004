"""N > 1 path on CPU: world_size-2 gloo processes partition independent
sequences (sequence s -> rank s mod G), track them (here with the oracle as the
stand-in engine -- no GPU in this container), gather the tables on rank 0 and
reduce the throughput counters.  The gathered result must not depend on G."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "klt-feature-tracker-acceleration-gpus_b200"
NSEQ, NFRAMES, NFEAT, W, H = 5, 3, 40, 160, 120


def _track_one(seq):
    """one independent sequence on the CPU oracle"""
    import importlib
    sys.path.insert(0, ROOT)
    synth = importlib.import_module(PKG + ".synth")
    from oracle import oracle_py
    o = oracle_py.Oracle()
    p = o.default_params()
    rng = np.random.default_rng(1000 + seq)
    vel = rng.uniform(-3, 3, 2)
    frames = [synth.frame(W, H, seed=1000 + seq, t=float(t), velocity=vel, threads=1) for t in range(NFRAMES)]
    tab = np.zeros((NFEAT, NFRAMES), dtype=[("x", "f4"), ("y", "f4"), ("val", "i4")])
    x, y, v = o.select(frames[0], p, NFEAT)
    tab["x"][:, 0], tab["y"][:, 0], tab["val"][:, 0] = x, y, v
    prev = o.build_pyramids(frames[0], p)
    for i in range(1, NFRAMES):
        cur = o.build_pyramids(frames[i], p)
        x, y, v = o.track(prev, cur, p, x, y, v)
        tab["x"][:, i], tab["y"][:, i], tab["val"][:, i] = x, y, v
        prev = cur
    return tab


def _worker(rank, world, port, q):
    import importlib
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    ms = importlib.import_module(PKG + ".multiseq")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = ms.my_sequences(NSEQ, rank, world)
    local = ms.run_shard(mine, _track_one)
    feats = float(sum((t["val"][:, :-1] >= 0).sum() for t in local.values()))
    total, tmax = ms.aggregate(feats, 1.0 + rank, world, dist)
    tables = ms.gather_tables(local, rank, world, dist)
    if rank == 0:
        q.put((total, tmax, {k: v.tobytes() for k, v in tables.items()}, mine))
    dist.barrier()
    dist.destroy_process_group()


def test_assignment_is_round_robin(pkg):
    import importlib
    ms = importlib.import_module(pkg.__name__ + ".multiseq")
    assert ms.assign(64, 8)[3] == list(range(3, 64, 8))
    for g in (1, 2, 4, 8):
        parts = ms.assign(64, g)
        assert sorted(sum(parts, [])) == list(range(64))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert ms.assign(3, 4) == [[0], [1], [2], []]          # ragged: an idle rank


def test_two_rank_gloo_matches_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, tmax, tables, mine0 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert mine0 == [0, 2, 4]
    assert tmax == 2.0                                      # max over ranks
    assert sorted(tables) == list(range(NSEQ))
    want = {s: _track_one(s) for s in range(NSEQ)}
    assert total == float(sum((t["val"][:, :-1] >= 0).sum() for t in want.values()))
    for s in range(NSEQ):
        assert tables[s] == want[s].tobytes(), "sequence %d differs between G=1 and G=2" % s
