"""GPU: k_check_solved (through the C ABI, via diffusion_ccsp_b200.checker.SolvedChecker) against the CPU oracle
(oracle/checker_oracle.py, itself pinned to the reference's labeller by tests/test_checker_oracle.py).

Integer / boolean work: the solved flags and the (collision, missing-constraint) counts must match EXACTLY.
"""
import numpy as np
import pytest
import torch

from diffusion_ccsp_b200 import scenes, synthetic
from diffusion_ccsp_b200.checker import SolvedChecker
from oracle import checker_oracle as chk
from oracle import ref_shim

pytestmark = pytest.mark.gpu

QDIMS = synthetic.DIMS['qualitative']
BDIMS = synthetic.DIMS['diffuse_pairwise']


def gpu_check(batch, poses, dims, mode):
    ck = SolvedChecker(batch, dims, mode, 'cuda')
    solved, counts = ck(torch.from_numpy(np.ascontiguousarray(poses, dtype=np.float32)), return_counts=True)
    return solved.cpu().numpy(), counts.cpu().numpy()


def assert_same(batch, poses, dims, mode, labeller=None):
    qual = 'qualitative' in mode
    s_ref, c_ref, m_ref = chk.check_batch(poses, batch, (dims[-1][1], dims[-1][2]), qualitative=qual, labeller=labeller)
    s, cnt = gpu_check(batch, poses, dims, mode)
    assert np.array_equal(s, s_ref), np.where(s != s_ref)[0][:10]
    assert np.array_equal(cnt[:, 0], c_ref) and np.array_equal(cnt[:, 1], m_ref), \
        (np.where(cnt[:, 0] != c_ref)[0][:10], np.where(cnt[:, 1] != m_ref)[0][:10])
    return s


def perturbed(batch, dims, rng, sigma):
    """ground-truth poses with Gaussian jitter: a mix of solved, colliding and relation-violating scenes"""
    p0, p1 = dims[-1][1], dims[-1][2]
    gt = batch.x[:, p0:p1].numpy().copy()
    free = ~batch.mask.numpy().astype(bool)
    out = gt.copy()
    out[free] += rng.normal(0, sigma, out[free].shape).astype(np.float32)
    return out


@pytest.mark.parametrize('n_obj', [3, 4, 6, 8])
def test_ground_truth_layouts_are_solved(n_obj):
    batch = scenes.qualitative_batch(16 if n_obj in (3, 6) else 64, n_obj)
    s = assert_same(batch, batch.x[:, 2:6].numpy(), QDIMS, 'qualitative')
    assert s.all()


@pytest.mark.parametrize('sigma', [1e-4, 3e-3, 0.02, 0.1, 0.5])
def test_jittered_layouts_match_oracle_exactly(sigma):
    rng = np.random.default_rng(int(sigma * 1e5))
    batch = scenes.qualitative_batch(256, 8, seed=3)
    poses = perturbed(batch, QDIMS, rng, sigma)
    s = assert_same(batch, poses, QDIMS, 'qualitative')
    if sigma <= 1e-4:
        assert s.mean() > 0.5
    if sigma >= 0.5:
        assert s.mean() < 0.1


def test_against_the_reference_labeller_when_available():
    """same comparison with the reference's OWN compute_qualitative_constraints doing the re-derivation"""
    if not ref_shim.reference_available():
        pytest.skip('reference sources not staged')
    _, du = ref_shim.load_reference_envs()
    rng = np.random.default_rng(1)
    batch = scenes.qualitative_batch(128, 8, seed=5)
    for sigma in (1e-3, 0.03):
        assert_same(batch, perturbed(batch, QDIMS, rng, sigma), QDIMS, 'qualitative', labeller=du.compute_qualitative_constraints)


def test_rotated_tiles_and_arbitrary_yaw():
    """free (cs, sn) columns: the sampler's output is not on the unit circle; yaw = atan2 after normalisation"""
    rng = np.random.default_rng(2)
    batch = scenes.qualitative_batch(128, 4, seed=2)
    poses = perturbed(batch, QDIMS, rng, 0.01)
    free = ~batch.mask.numpy().astype(bool)
    poses[free, 2:] = rng.uniform(-1.3, 1.3, poses[free, 2:].shape).astype(np.float32)      # also exercises the clamp
    assert_same(batch, poses, QDIMS, 'qualitative')


def test_nan_rows_clamp_and_out_of_tray():
    batch = scenes.qualitative_batch(32, 4)
    poses = batch.x[:, 2:6].numpy().copy()
    poses[1, 0] = np.nan                # scene 0: NaN -> skipped = unsolved (ddpm.py:644-645)
    poses[7, 1] = 7.5                   # scene 1: clamped to 1 -> sticks out of the tray
    poses[11:15, :2] = 0.0              # scene 2: all tiles on one spot
    s, cnt = gpu_check(batch, poses, QDIMS, 'qualitative')
    s_ref, c_ref, m_ref = chk.check_batch(poses, batch, (2, 6))
    assert np.array_equal(s, s_ref) and not s[0] and not s[1] and not s[2] and s[3:].all()
    assert tuple(cnt[0]) == (-1, -1) and cnt[1, 0] > 0 and cnt[2, 0] > 0
    assert np.array_equal(cnt[1:, 0], c_ref[1:]) and np.array_equal(cnt[1:, 1], m_ref[1:])


def test_boxes_world_collisions_only():
    rng = np.random.default_rng(3)
    batch = scenes.make_batch('boxes', 128, 12, seed=7)
    gt = batch.x[:, 2:4].numpy()
    s = assert_same(batch, gt, BDIMS, 'diffuse_pairwise')
    assert s.all()
    for sigma in (0.01, 0.05, 0.3):
        assert_same(batch, perturbed(batch, BDIMS, rng, sigma), BDIMS, 'diffuse_pairwise')


def test_ragged_scenes_and_foreign_edge_types():
    """scenes of different sizes in one batch, a scene with out-of-vocabulary type ids (skipped, data_utils.py:180-181) and
    constraints that name a relation the layout does not have"""
    rng = np.random.default_rng(4)
    parts = [scenes.qualitative_batch(4, n, seed=n).select_scenes(i, i + 1) for n in (3, 4, 6, 8) for i in range(4)]
    batch = scenes.collate(parts)
    ea = batch.edge_attr.clone()
    ea[::17] = 13.0 + (torch.arange(ea[::17].numel()) % 3).float()          # unknown ids: ignored by the check
    batch = scenes.SceneBatch(batch.x, batch.edge_index, ea, batch.mask)
    assert_same(batch, batch.x[:, 2:6].numpy(), QDIMS, 'qualitative')
    # flip relation types at random: most scenes now hold a constraint that the layout does not satisfy
    ea2 = batch.edge_attr.clone()
    idx = rng.choice(ea2.numel(), ea2.numel() // 5, replace=False)
    ea2[idx] = torch.from_numpy(rng.integers(0, 13, idx.size).astype(np.float32))
    b2 = scenes.SceneBatch(batch.x, batch.edge_index, ea2, batch.mask)
    s = assert_same(b2, b2.x[:, 2:6].numpy(), QDIMS, 'qualitative')
    assert not s.all()


def test_full_batch_sampler_output_is_checked_in_one_launch():
    """config-2 sized batch: sampler output -> checker, both on the device; flags equal the oracle's"""
    from diffusion_ccsp_b200 import _abi
    from diffusion_ccsp_b200.ddpm import GaussianDiffusion
    from diffusion_ccsp_b200.denoise_fn import ConstraintDiffuser
    batch = scenes.qualitative_batch(1024, 8)
    den = ConstraintDiffuser(dims=QDIMS, input_mode='qualitative', device='cuda', verbose=False, math='bf16x3')
    gd = GaussianDiffusion(den, timesteps=20, EBM='ULA', samples_per_step=4).eval()
    gd.load_state_dict(synthetic.make_trained_state_dict(), strict=False)
    poses = gd.sample(batch, seed=7)
    ck = SolvedChecker(batch, QDIMS, 'qualitative', 'cuda')
    _abi.reset_launch_count()
    solved, counts = ck(poses, return_counts=True)
    assert _abi.launch_count() == 1
    s_ref, c_ref, m_ref = chk.check_batch(poses.cpu().numpy(), batch, (2, 6))
    assert np.array_equal(solved.cpu().numpy(), s_ref)
    assert np.array_equal(counts.cpu().numpy()[:, 0], c_ref) and np.array_equal(counts.cpu().numpy()[:, 1], m_ref)


def test_scene_and_edge_order_do_not_matter():
    """per-scene verdicts and counts are invariant under the position of the scene in the batch and under the order of its edges"""
    rng = np.random.default_rng(9)
    batch = scenes.qualitative_batch(96, 8, seed=4)
    poses = perturbed(batch, QDIMS, rng, 0.004)
    s0, c0 = gpu_check(batch, poses, QDIMS, 'qualitative')
    # scenes in another order
    order = rng.permutation(batch.num_graphs)
    off = batch.scene_node_ranges()
    b2 = scenes.take_scenes(batch, order)
    p2 = np.concatenate([poses[off[i]:off[i + 1]] for i in order])
    s2, c2 = gpu_check(b2, p2, QDIMS, 'qualitative')
    assert np.array_equal(s2, s0[order]) and np.array_equal(c2, c0[order])
    # edges in another order (the scene id of every edge travels with it)
    perm = rng.permutation(batch.num_edges)
    b3 = scenes.SceneBatch(batch.x, batch.edge_index[:, perm], batch.edge_attr[perm], batch.mask, batch.x_extract, batch.edge_extract[perm])
    s3, c3 = gpu_check(b3, poses, QDIMS, 'qualitative')
    assert np.array_equal(s3, s0) and np.array_equal(c3, c0)
    assert 0 < s0.sum() < s0.size
