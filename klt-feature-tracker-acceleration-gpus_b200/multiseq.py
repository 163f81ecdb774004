"""Multi-sequence partitioning (BASELINE config 5).

Independent image sequences share nothing: sequence s runs on rank s mod G with
its own tracking context, stream and pyramids; there is NO collective on the
data path.  The only communication is the host-side gather of the small feature
tables (12 B x nFeatures per frame) and the (sum, max) reduction of the
throughput counters -- both through torch.distributed (NCCL on the GPU box,
gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np


def assign(nseq: int, world: int) -> List[List[int]]:
    """Round-robin: sequence s -> rank s mod world."""
    return [[s for s in range(nseq) if s % world == r] for r in range(world)]


def my_sequences(nseq: int, rank: int, world: int) -> List[int]:
    return assign(nseq, world)[rank]


def run_shard(seq_ids: Sequence[int], track_one: Callable[[int], np.ndarray]) -> Dict[int, np.ndarray]:
    """Run this rank's sequences one after the other; track_one(s) returns the
    feature table of sequence s as a structured array [nFeatures, nFrames]."""
    return {int(s): track_one(int(s)) for s in seq_ids}


def gather_tables(local: Dict[int, np.ndarray], rank: int, world: int, dist=None) -> Dict[int, np.ndarray]:
    """Host gather of the per-sequence tables onto rank 0 (others get {})."""
    if world == 1 or dist is None:
        return dict(local)
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(local, bucket, dst=0)
    if rank != 0:
        return {}
    out: Dict[int, np.ndarray] = {}
    for part in bucket:
        out.update(part)
    return out


def aggregate(features: float, seconds: float, world: int, dist=None, device=None) -> Tuple[float, float]:
    """(sum over ranks of features, max over ranks of seconds)."""
    if world == 1 or dist is None:
        return float(features), float(seconds)
    import torch
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    c = torch.tensor([float(features)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(c[0]), float(t[0])
