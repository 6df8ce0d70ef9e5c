"""Build-container only: pin oracle/ccsp_oracle.py against the UNMODIFIED reference on inputs the committed
goldens do not contain (random typed graphs, other seeds).  Skipped where /root/reference is absent
(e.g. on the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import ccsp_oracle as orc
from oracle.ref_shim import injected_randn, load_reference, reference_available
from diffusion_ccsp_b200 import scenes, synthetic
from tests.util import rel_err

pytestmark = [pytest.mark.reference, pytest.mark.skipif(not reference_available(), reason='needs /root/reference')]


def ref_models(mode, dims, sd, T, EBM, K):
    dfn, ddpm = load_reference()
    m = dfn.ConstraintDiffuser(dims=dims, hidden_dim=256, input_mode=mode, EBM=EBM, device='cpu', verbose=False)
    gd = ddpm.GaussianDiffusion(m, timesteps=T, EBM=EBM, samples_per_step=K if K else 10).eval()
    gd.load_state_dict(sd, strict=False)
    return m, gd


@pytest.mark.parametrize('seed', [3, 4])
def test_forward_random_typed_graph(seed):
    rng = np.random.default_rng(seed)
    mode, dims = 'qualitative', synthetic.DIMS['qualitative']
    sd = synthetic.make_state_dict(dims, mode, seed=seed)
    batch = scenes.collate([scenes.random_typed_scene(rng, int(rng.integers(2, 9)), 13, int(rng.integers(5, 40)), 6)
                            for _ in range(5)])
    poses = rng.standard_normal((batch.num_nodes, 4)).astype(np.float32)
    m, _ = ref_models(mode, dims, sd, 50, 'ULA', 10)
    den = orc.OracleDenoiser({k: v.numpy() for k, v in sd.items()}, dims, mode)
    for t in (0, 21, 49):
        with torch.no_grad():
            ref = m(torch.from_numpy(poses.copy()), batch, torch.tensor([t]), eval=True).detach().numpy()
        torch.set_grad_enabled(True)
        assert rel_err(den.forward(poses, batch, t), ref) < 2e-6


@pytest.mark.parametrize('case', [('boxes', 'diffuse_pairwise', False, 6), ('robot_box', 'robot_box', False, 4)])
def test_short_trajectory_other_seed(case):
    kind, mode, tri, n_obj = case
    dims = synthetic.dims_for(mode, tri)
    sd = synthetic.make_state_dict(dims, mode, seed=31)
    batch = scenes.make_batch(kind, 4, n_obj, seed=9)
    T, K = 4, 2
    noise = synthetic.make_noise(T, K, batch.num_nodes, dims[-1][0], seed=77)
    _, gd = ref_models(mode, dims, sd, T, 'ULA', K)
    with injected_randn(noise) as inj:
        ref = gd.sample(batch).detach().numpy()
    assert inj.calls == orc.num_noise_draws(T, K)
    den = orc.OracleDenoiser({k: v.numpy() for k, v in sd.items()}, dims, mode)
    out = orc.OracleDiffusion(den, T, 'ULA', K).p_sample_loop(batch, noise.numpy())
    assert rel_err(out, ref) < 1e-5
