"""Dataset normalisation statistics (H12 of SURVEY.md section 8a; north_star config 4).

Per mel bin: sum, sum of squares and frame count over every REAL (un-padded, un-masked,
un-normalised) frame, accumulated in float64 inside the fused kernel
(``b200fbank_stats_accumulate``).  Clips are sharded by batch across ranks; the only
collective of the whole path is one all-reduce of the ``2*n_cols+1`` doubles.
"""
from __future__ import annotations

import math
from typing import NamedTuple, Optional

import torch
import torch.distributed as dist

from .frontend import FbankFrontend


class NormStats(NamedTuple):
    mean_per_bin: torch.Tensor   # float64 [n_cols]
    std_per_bin: torch.Tensor    # float64 [n_cols]   (population)
    mean: float
    std: float
    frames: int


def finalize_sums(sums: torch.Tensor) -> NormStats:
    """Population convention: mean_b = S/N, std_b = sqrt(SS/N - mean_b^2); the scalars pool all bins."""
    s = sums.detach().to("cpu", torch.float64)
    n = (s.numel() - 1) // 2
    cnt = float(s[2 * n])
    if cnt <= 0:
        raise ValueError("no frames accumulated")
    mean_b = s[:n] / cnt
    std_b = torch.sqrt(torch.clamp(s[n:2 * n] / cnt - mean_b * mean_b, min=0.0))
    gm = float(s[:n].sum() / (cnt * n))
    gs = math.sqrt(max(float(s[n:2 * n].sum() / (cnt * n)) - gm * gm, 0.0))
    return NormStats(mean_b, std_b, gm, gs, int(cnt))


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous batch-index shard of rank ``rank``."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


def shard_by_samples(lengths, rank: int, world: int):
    """Ragged batches: contiguous shards balanced on cumulative SAMPLE count, not clip count."""
    lengths = [int(x) for x in lengths]
    total = sum(lengths)
    bounds, acc, r = [0], 0, 1
    for i, n in enumerate(lengths):
        acc += n
        while r < world and acc >= total * r / world:
            bounds.append(i + 1)
            r += 1
    while len(bounds) < world + 1:
        bounds.append(len(lengths))
    bounds[-1] = len(lengths)
    return bounds[rank], bounds[rank + 1]


class DatasetStats:
    """Accumulate on one GPU, all-reduce across ranks, finalise."""

    def __init__(self, frontend: FbankFrontend, max_frames: int):
        self.fe = frontend
        self.max_frames = int(max_frames)
        self.sums = torch.zeros(2 * frontend.n_cols + 1, dtype=torch.float64, device=frontend.device)

    def update(self, wav: torch.Tensor, offsets: Optional[torch.Tensor] = None,
               rate_ids: Optional[torch.Tensor] = None) -> "DatasetStats":
        self.fe.accumulate_stats(wav, self.sums, self.max_frames, offsets=offsets, rate_ids=rate_ids)
        return self

    def all_reduce(self, group=None) -> "DatasetStats":
        """The path's single collective: SUM over ranks of 2*n_cols+1 float64 (about 2 KB, latency-bound)."""
        all_reduce_sums(self.sums, group)
        return self

    def finalize(self) -> NormStats:
        return finalize_sums(self.sums)


def all_reduce_sums(sums: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums
