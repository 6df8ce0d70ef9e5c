"""GPU: the opt-in persistent small-shard path (CCSP_PERSIST=1: two kernels launched once per sample(), hand-shaking through device flags) gives
bit-identical results to the launch-per-evaluation path — same arithmetic, same order — for Philox and injected noise, with
history, for every pose width, and falls back by itself when the shard does not fit."""
import os

import numpy as np
import pytest
import torch

from diffusion_ccsp_b200 import _abi, scenes, synthetic
from diffusion_ccsp_b200.ddpm import GaussianDiffusion
from diffusion_ccsp_b200.denoise_fn import ConstraintDiffuser

pytestmark = pytest.mark.gpu


def same_bits(a, b):
    """bit-for-bit equality (NaNs included: short schedules with trained weights can overflow)"""
    return torch.equal(a.contiguous().view(torch.int32), b.contiguous().view(torch.int32))


def run(batch, mode, tri, T, K, persist, EBM='ULA', trained=False, **kw):
    dims = synthetic.dims_for(mode, tri)
    sd = synthetic.load_trained_checkpoint() if trained else synthetic.make_state_dict(dims, mode, seed=3)
    den = ConstraintDiffuser(dims=dims, input_mode=mode, device='cuda', verbose=False, math='bf16x3')
    gd = GaussianDiffusion(den, timesteps=T, EBM=EBM, samples_per_step=K if K else 10).eval()
    gd.load_state_dict(sd, strict=False)
    os.environ['CCSP_PERSIST'] = '1' if persist else '0'
    try:
        den.plan_for(batch)                     # the plan's own launches are not what is counted
        _abi.reset_launch_count()
        out = gd.sample(batch, **kw)
        torch.cuda.synchronize()
        return out, _abi.launch_count()
    finally:
        os.environ.pop('CCSP_PERSIST', None)


CASES = [('qualitative', False, lambda: scenes.qualitative_batch(8, 4), True),            # config 1
         ('qualitative', False, lambda: scenes.qualitative_batch(128, 8), True),          # the N = 8 strong-scaling shard
         ('diffuse_pairwise', False, lambda: scenes.make_batch('boxes', 64, 12, seed=1), False),
         ('diffuse_pairwise', True, lambda: scenes.make_batch('triangles', 96, 10, seed=2), False),
         ('robot_box', False, lambda: scenes.make_batch('robot_box', 256, 6, seed=3), False)]   # config 5 shard


@pytest.mark.parametrize('case', CASES, ids=['config1', 'n8_shard128', 'boxes64', 'triangles96', 'robot256'])
def test_persistent_equals_launch_per_evaluation(case):
    mode, tri, factory, trained = case
    batch = factory()
    T, K = 12, 5
    a, la = run(batch, mode, tri, T, K, persist=False, trained=trained, seed=11)
    b, lb = run(batch, mode, tri, T, K, persist=True, trained=trained, seed=11)
    # (+2 when the model's time table is built inside the call)
    assert la - (1 + 2 * T * (1 + K)) in (0, 2) and lb in (2, 4), (la, lb)     # the persistent path really ran: two launches per sample
    assert same_bits(a, b)
    # a second sample on the same plan (flags are reset, barriers re-initialised by the fresh launch)
    c, _ = run(batch, mode, tri, T, K, persist=True, trained=trained, seed=12)
    d, _ = run(batch, mode, tri, T, K, persist=False, trained=trained, seed=12)
    assert same_bits(c, d) and not same_bits(a, c)


def test_persistent_history_injected_noise_ddpm_and_ulaplus():
    batch = scenes.qualitative_batch(16, 4)
    dims = synthetic.DIMS['qualitative']
    for EBM, T, K in (('ULA', 6, 3), (False, 9, 0), ('ULA+', 8, 10)):
        draws = 1 + T + (2 * (4 + 8 + 12 + 16) if EBM == 'ULA+' else T * K)
        noise = torch.from_numpy(np.random.default_rng(T).standard_normal((draws, batch.num_nodes, 4), dtype=np.float32))
        (a, ha), _ = run(batch, 'qualitative', False, T, K, persist=False, EBM=EBM, trained=True, noise=noise, return_history=True)
        (b, hb), lb = run(batch, 'qualitative', False, T, K, persist=True, EBM=EBM, trained=True, noise=noise, return_history=True)
        assert lb in (2, 4) and same_bits(a, b)
        assert len(ha) == len(hb) == T + 1 and all(same_bits(x, y) for x, y in zip(ha, hb))


def test_large_shards_fall_back_to_the_launch_per_evaluation_path():
    batch = scenes.qualitative_batch(1024, 8)
    out, launches = run(batch, 'qualitative', False, 2, 2, persist=True, trained=True, seed=5)
    assert launches - (1 + 2 * 2 * 3) in (0, 2)                       # 645 units do not fit next to 145 node CTAs
    assert bool(torch.isfinite(out).all())


# ---- pipelined chains (opt-in, CCSP_CHAINS=2..4 at plan creation): the plan is cut into independent scene groups, the persistent
# edge kernel walks (evaluation, chain) pairs while a few persistent node CTAs serve the node blocks of the other chain -------------
def run_chains(batch, mode, tri, T, K, chains, node_ctas=None, EBM='ULA', trained=False, **kw):
    dims = synthetic.dims_for(mode, tri)
    sd = synthetic.load_trained_checkpoint() if trained else synthetic.make_state_dict(dims, mode, seed=3)
    den = ConstraintDiffuser(dims=dims, input_mode=mode, device='cuda', verbose=False, math='bf16x3')
    gd = GaussianDiffusion(den, timesteps=T, EBM=EBM, samples_per_step=K if K else 10).eval()
    gd.load_state_dict(sd, strict=False)
    os.environ['CCSP_CHAINS'] = str(chains)
    if node_ctas:
        os.environ['CCSP_PIPE_NODE_CTAS'] = str(node_ctas)
    try:
        den.plan_for(batch)
        _abi.reset_launch_count()
        out = gd.sample(batch, **kw)
        torch.cuda.synchronize()
        return out, _abi.launch_count()
    finally:
        os.environ.pop('CCSP_CHAINS', None)
        os.environ.pop('CCSP_PIPE_NODE_CTAS', None)


@pytest.mark.parametrize('chains,node_ctas', [(2, 8), (3, 24), (4, 2), (2, None), (4, None)])     # None: one node CTA per block and chain
def test_pipelined_chains_equal_launch_per_evaluation(chains, node_ctas):
    batch = scenes.qualitative_batch(200, 8)                  # 1800 nodes, ~15.8 k edges: several units per pair and chain
    T, K = 8, 4
    a, la = run_chains(batch, 'qualitative', False, T, K, 1, trained=True, seed=21)
    b, lb = run_chains(batch, 'qualitative', False, T, K, chains, node_ctas, trained=True, seed=21)
    assert la - (1 + 2 * T * (1 + K)) in (0, 2) and lb in (2, 4), (la, lb)      # two launches per sample
    assert same_bits(a, b)
    c, _ = run_chains(batch, 'qualitative', False, T, K, chains, node_ctas, trained=True, seed=22)
    assert not same_bits(b, c)


def test_pipelined_chains_history_noise_and_other_worlds():
    batch = scenes.qualitative_batch(48, 4)
    T, K = 5, 3
    noise = torch.from_numpy(np.random.default_rng(1).standard_normal((1 + T * (1 + K), batch.num_nodes, 4), dtype=np.float32))
    (a, ha), _ = run_chains(batch, 'qualitative', False, T, K, 1, trained=True, noise=noise, return_history=True)
    (b, hb), lb = run_chains(batch, 'qualitative', False, T, K, 2, 4, trained=True, noise=noise, return_history=True)
    assert lb in (2, 4) and same_bits(a, b) and all(same_bits(x, y) for x, y in zip(ha, hb))
    for mode, tri, factory in (('diffuse_pairwise', False, lambda: scenes.make_batch('boxes', 64, 12, seed=1)),
                               ('robot_box', False, lambda: scenes.make_batch('robot_box', 96, 6, seed=3))):
        bt = factory()
        a, _ = run_chains(bt, mode, tri, 6, 3, 1, seed=5)
        b, lb = run_chains(bt, mode, tri, 6, 3, 3, 6, seed=5)
        assert lb in (2, 4) and same_bits(a, b)


def test_chain_cut_needs_independent_scene_groups():
    """one scene cannot be cut: the plan stays a single chain and sampling takes the default path"""
    batch = scenes.qualitative_batch(1, 4)
    out, launches = run_chains(batch, 'qualitative', False, 3, 2, 2, trained=True, seed=1)
    assert launches - (1 + 2 * 3 * 3) in (0, 2) and bool(torch.isfinite(out).all())


def test_pipelined_chains_report_kernel_timing():
    """bench.py's roofline leg on the pipelined path: the persistent edge kernel is bracketed by one event pair and booked as
    `evaluations` samples, so avg = duration / evaluations"""
    batch = scenes.qualitative_batch(200, 8)
    dims = synthetic.DIMS['qualitative']
    den = ConstraintDiffuser(dims=dims, input_mode='qualitative', device='cuda', verbose=False, math='bf16x3')
    T, K = 6, 4
    gd = GaussianDiffusion(den, timesteps=T, EBM='ULA', samples_per_step=K).eval()
    gd.load_state_dict(synthetic.load_trained_checkpoint(), strict=False)
    os.environ['CCSP_CHAINS'] = '2'
    try:
        plan = den.plan_for(batch)
        plan.set_timing(1)
        gd.sample(batch, seed=3)
        gd.sample(batch, seed=4)
        tm = plan.get_timing()
        plan.set_timing(0)
    finally:
        os.environ.pop('CCSP_CHAINS', None)
    assert tm['samples'] == 2 * T * (1 + K)
    assert 0.0 < tm['ms_edge_l1'] / tm['samples'] < 1.0 and tm['ms_node'] == 0.0
