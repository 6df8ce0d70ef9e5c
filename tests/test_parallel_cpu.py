"""CPU (gloo, world_size 2): the scene-sharding / gather plumbing of diffusion_ccsp_b200.parallel.
The CUDA sampler is replaced by a deterministic CPU function of (features, global node id) so the test
checks exactly what the multi-GPU path adds: shard boundaries, node re-basing, noise slicing, node_offset,
and the padded all-gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffusion_ccsp_b200 import parallel, scenes


def fake_sampler(local, node_offset, local_noise):
    # depends on node features, on re-based edge structure, on the global node id and on the noise slice
    n = local.num_nodes
    deg = torch.zeros(n).index_add_(0, local.edge_index.reshape(-1), torch.ones(local.edge_index.numel()))
    gid = torch.arange(n, dtype=torch.float32) + node_offset
    out = torch.stack([local.x[:, 2], local.x[:, 3], deg, gid], 1)
    if local_noise is not None:
        out = out + local_noise.sum(0)
    return out


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        batch = scenes.collate([scenes.qualitative_batch(5, 4), scenes.qualitative_batch(4, 6), scenes.qualitative_batch(2, 3)])
        noise = torch.from_numpy(np.random.default_rng(0).standard_normal((3, batch.num_nodes, 4), dtype=np.float32))
        full = parallel.sample_sharded(None, batch, noise=noise, sampler=fake_sampler)
        scn, sec = parallel.reduce_run_stats(parallel.shard_bounds(batch.num_graphs, rank, world)[1]
                                             - parallel.shard_bounds(batch.num_graphs, rank, world)[0], 1.0 + rank, 'cpu')
        ret[rank] = (full.numpy(), scn, sec)
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for n in (1, 7, 8, 1024, 1025):
        for w in (1, 2, 3, 8):
            b = [scenes.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1


def test_select_scenes_rebases_edges():
    batch = scenes.qualitative_batch(8, 4)
    sub = batch.select_scenes(3, 6)
    assert sub.num_graphs == 3 and sub.num_nodes == 15
    assert int(sub.edge_index.min()) >= 0 and int(sub.edge_index.max()) < 15
    assert sub.mask.sum() == 3 and bool(sub.mask[0])
    off = batch.scene_node_ranges()
    assert torch.equal(sub.x, batch.x[off[3]:off[6]])


@pytest.mark.timeout(120)
def test_sharded_sampling_matches_single_process_gloo():
    batch = scenes.collate([scenes.qualitative_batch(5, 4), scenes.qualitative_batch(4, 6), scenes.qualitative_batch(2, 3)])
    noise = torch.from_numpy(np.random.default_rng(0).standard_normal((3, batch.num_nodes, 4), dtype=np.float32))
    single = fake_sampler(batch, 0, noise).numpy()
    # the fake sampler's degree term must be computed per shard on re-based edges and still agree
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        full, scn, sec = ret[r]
        assert full.shape == single.shape
        assert np.array_equal(full, single), f'rank {r} gathered result differs from the single-process run'
        assert scn == batch.num_graphs and sec == 2.0
