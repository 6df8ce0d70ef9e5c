"""Scene-sharded multi-GPU sampling.

The sampling path is embarrassingly parallel over scenes: a collated batch is a disjoint union of scene
graphs, weights/schedule are replicated, and the in-kernel Philox stream is keyed on the GLOBAL node id.
So there is no collective inside the T x (1+K) loop; each rank samples a contiguous shard of scenes and
the final poses are gathered once at the end (one `all_gather_into_tensor` of fixed-size padded shards
over NCCL/NVLink — a few hundred KB, latency-bound).  One process per GPU (torchrun).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch
import torch.distributed as dist

from .scenes import SceneBatch, shard_bounds


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_layout(batch: SceneBatch, world_size: int):
    """node ranges [(n0, n1)] of every rank's contiguous scene shard."""
    off = batch.scene_node_ranges()
    out = []
    for r in range(world_size):
        lo, hi = shard_bounds(batch.num_graphs, r, world_size)
        out.append((int(off[lo]), int(off[hi])))
    return out


class ShardedSampler:
    """One global batch, its scenes sharded over the ranks of `group`: the shard is cut (and its plan compiled) once, every
    `sample()` runs the local T x (1+K) loop with no collective inside and ends with ONE `all_gather_into_tensor` of the final
    poses (fixed-size padded shards).  With `seed` the result is bit-identical to the single-GPU run for any world size (Philox
    is keyed on the global node id); with `noise` [draws, n, P] each rank consumes its node slice."""

    def __init__(self, diffusion, batch: SceneBatch, group=None, sampler: Optional[Callable] = None):
        self.diffusion, self.batch, self.group, self.sampler = diffusion, batch, group, sampler
        self.rank, self.world = _world(group)
        self.layout = shard_layout(batch, self.world)
        self.n0, self.n1 = self.layout[self.rank]
        lo, hi = shard_bounds(batch.num_graphs, self.rank, self.world)
        self.scene_range = (lo, hi)
        self.local = batch.select_scenes(lo, hi) if self.world > 1 else batch       # kept: the plan cache keys on these tensors
        self.max_n = max(b - a for a, b in self.layout)
        self._padded = self._gathered = None

    def sample_local(self, *, seed: Optional[int] = None, noise: Optional[torch.Tensor] = None, **kwargs) -> torch.Tensor:
        local_noise = noise[:, self.n0:self.n1] if noise is not None else None
        if self.sampler is not None:
            return self.sampler(self.local, self.n0, local_noise)
        return self.diffusion.p_sample_loop(self.local, seed=seed, noise=local_noise, node_offset=self.n0, **kwargs)

    def gather(self, out: torch.Tensor) -> torch.Tensor:
        """poses of ALL nodes [n, P] on every rank, same row order as the global batch"""
        if self.world == 1:
            return out
        P = out.shape[1]
        if self._padded is None or self._padded.shape[1] != P or self._padded.device != out.device:
            self._padded = torch.zeros((self.max_n, P), dtype=out.dtype, device=out.device)
            self._gathered = torch.empty((self.world * self.max_n, P), dtype=out.dtype, device=out.device)
        self._padded[: self.n1 - self.n0] = out
        dist.all_gather_into_tensor(self._gathered, self._padded, group=self.group)
        if all(b - a == self.max_n for a, b in self.layout):
            return self._gathered
        return torch.cat([self._gathered[r * self.max_n: r * self.max_n + (b - a)] for r, (a, b) in enumerate(self.layout)], 0)

    def sample(self, **kwargs) -> torch.Tensor:
        return self.gather(self.sample_local(**kwargs))


def sample_sharded(diffusion, batch: SceneBatch, *, seed: Optional[int] = None, noise: Optional[torch.Tensor] = None,
                   group=None, sampler: Optional[Callable] = None, **kwargs) -> torch.Tensor:
    """Sample `batch` with its scenes sharded over the ranks of `group`; returns the poses of ALL nodes [n, P] on every rank
    (same row order as `batch`).  One-shot form of `ShardedSampler` (which keeps the shard and its plan across calls).
    `sampler(local_batch, node_offset, local_noise)` can replace the CUDA sampler (used by the CPU/gloo tests of this plumbing)."""
    return ShardedSampler(diffusion, batch, group, sampler).sample(seed=seed, noise=noise, **kwargs)


def reduce_run_stats(num_scenes: int, seconds: float, device, group=None):
    """(total scenes, max wall time over ranks): the end-of-run counters of SURVEY.md §8e."""
    rank, world = _world(group)
    if world == 1:
        return num_scenes, seconds
    s = torch.tensor([float(num_scenes)], dtype=torch.float64, device=device)
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(s.item()), float(t.item())
