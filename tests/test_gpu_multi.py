"""GPU, >= 2 devices: the product's sharded sampling path (parallel.ShardedSampler) over NCCL — two ranks, one process per GPU
(spawned with torch.multiprocessing; rendezvous on 127.0.0.1) — reproduces the single-GPU result bit for bit."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from diffusion_ccsp_b200 import scenes, synthetic
    from diffusion_ccsp_b200.ddpm import GaussianDiffusion
    from diffusion_ccsp_b200.denoise_fn import ConstraintDiffuser
    from diffusion_ccsp_b200.parallel import ShardedSampler, reduce_run_stats
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    dims = synthetic.DIMS['qualitative']
    batch = scenes.collate([scenes.qualitative_batch(37, 8), scenes.qualitative_batch(11, 4)])      # ragged: shards differ in size
    den = ConstraintDiffuser(dims=dims, input_mode='qualitative', device=dev, verbose=False, math='bf16x3')
    gd = GaussianDiffusion(den, timesteps=12, EBM='ULA', samples_per_step=4).eval()
    gd.load_state_dict(synthetic.load_trained_checkpoint(), strict=False)
    sampler = ShardedSampler(gd, batch)
    full = sampler.sample(seed=123)
    again = sampler.sample(seed=123)
    total, tmax = reduce_run_stats(sampler.local.num_graphs, 0.5 + rank, dev)
    single = gd.p_sample_loop(batch, seed=123) if rank == 0 else None
    q.put((rank, full.cpu(), torch.equal(full, again), total, tmax, single.cpu() if single is not None else None,
           sampler.scene_range))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_sharded_sampler_over_nccl_matches_single_gpu():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=600) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, full0, same0, total0, tmax0, single, range0), (r1, full1, same1, total1, tmax1, _, range1) = got
    assert same0 and same1
    assert torch.equal(full0, full1), 'every rank holds the same gathered poses'
    assert torch.equal(full0, single), 'sharded over 2 GPUs == single GPU, bit for bit (Philox keyed on the global node id)'
    assert total0 == total1 == 48 and tmax0 == tmax1 == 1.5
    assert range0 == (0, 24) and range1 == (24, 48)
