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


def sample_sharded(diffusion, batch: SceneBatch, *, seed: Optional[int] = None, noise: Optional[torch.Tensor] = None,
                   group=None, sampler: Optional[Callable] = None, **kwargs) -> torch.Tensor:
    """Sample `batch` with its scenes sharded over the ranks of `group`; returns the poses of ALL nodes
    [n, P] on every rank (same row order as `batch`).

    seed / noise follow GaussianDiffusion.p_sample_loop; with `seed` the result is bit-identical to the
    single-GPU run for any world size (Philox is keyed on the global node id), with `noise` [draws, n, P]
    each rank consumes its node slice.  `sampler(local_batch, node_offset, local_noise)` can replace the
    CUDA sampler (used by the CPU/gloo tests of this plumbing)."""
    rank, world = _world(group)
    layout = shard_layout(batch, world)
    n0, n1 = layout[rank]
    lo, hi = shard_bounds(batch.num_graphs, rank, world)
    local = batch.select_scenes(lo, hi) if world > 1 else batch
    local_noise = noise[:, n0:n1] if noise is not None else None
    if sampler is None:
        out = diffusion.p_sample_loop(local, seed=seed, noise=local_noise, node_offset=n0, **kwargs)
    else:
        out = sampler(local, n0, local_noise)
    if world == 1:
        return out
    P = out.shape[1]
    max_n = max(b - a for a, b in layout)
    padded = torch.zeros((max_n, P), dtype=out.dtype, device=out.device)
    padded[: n1 - n0] = out
    gathered = torch.empty((world * max_n, P), dtype=out.dtype, device=out.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    return torch.cat([gathered[r * max_n: r * max_n + (b - a)] for r, (a, b) in enumerate(layout)], 0)


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
