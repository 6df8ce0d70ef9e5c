"""GPU: the energy-form variant — ccsp_energy_grad through ComposedEBMDenoiseFn, and the ULA / MALA / HMC loops
around it — against golden vectors from the UNMODIFIED reference (tests/golden/energy_*.npz, ebm_*.npz; made by
make_energy_golden.py with the reference's autograd gradient).

Stated FP32 tolerance: energy 5e-6 relative, gradient 2e-5 of max|grad|, trajectories 2e-4 of max|x_ref| (untrained energies grow
the state to 1e9 over the 12 timesteps: a stress case; the accept / reject decisions must all agree for that to hold).
"""
import numpy as np
import pytest
import torch

from diffusion_ccsp_b200 import synthetic
from diffusion_ccsp_b200.ddpm import GaussianDiffusion
from diffusion_ccsp_b200.denoise_fn import ConstraintDiffuser
from diffusion_ccsp_b200.ebm import ComposedEBMDenoiseFn
from tests.util import case_model, golden_names, load_golden

pytestmark = pytest.mark.gpu


def build(mode, dims, sd, T, EBM, K, step_sizes='2*self.betas'):
    m = ConstraintDiffuser(dims=dims, input_mode=mode, EBM=EBM, energy_wrapper=True, device='cuda', verbose=False)
    w = ComposedEBMDenoiseFn(m, 1)
    gd = GaussianDiffusion(w, timesteps=T, EBM=EBM, samples_per_step=K, step_sizes=step_sizes).eval()
    sd = {k.replace('denoise_fn.', 'denoise_fn.model.'): v for k, v in sd.items()}
    missing, unexpected = gd.load_state_dict(sd, strict=False)
    assert not unexpected and not [k for k in missing if k.startswith('denoise_fn.')]
    m.to('cuda')
    return m, w, gd


@pytest.mark.parametrize('name', golden_names('energy_'))
def test_energy_and_gradient_vs_reference_autograd(name):
    z, batch = load_golden(name)
    mode, dims, sd = case_model(z)
    m, w, gd = build(mode, dims, sd, 100, 'ULA', 10)
    poses = torch.from_numpy(z['poses_in'])
    for t, g_ref, e_ref in zip(z['t'], z['grad'], z['energy']):
        g, e = m(poses, batch, torch.tensor([int(t)]), eval=True, tag='EBM')
        assert abs(float(e) - e_ref) / abs(e_ref) < 5e-6, (float(e), e_ref)
        err = float(np.max(np.abs(g.cpu().numpy() - g_ref)) / np.abs(g_ref).max())
        assert err < 2e-5, (name, t, err)
        assert torch.equal(w(poses, batch, torch.tensor([int(t)])), g)                 # wrapper: gradients only
        assert float(w.neg_logp_unnorm(poses, batch, torch.tensor([int(t)]))) == float(e)


@pytest.mark.parametrize('name', golden_names('ebm_'))
def test_energy_samplers_vs_reference_golden(name):
    z, batch = load_golden(name)
    mode, dims, sd = case_model(z)
    T, K, EBM = int(z['T']), int(z['K']), str(z['EBM'])
    m, w, gd = build(mode, dims, sd, T, EBM, K, step_sizes=str(z['step_sizes']))
    with torch.no_grad():
        m.pose_decoder[2].weight.mul_(float(z['decoder_scale'])); m.pose_decoder[2].bias.mul_(float(z['decoder_scale']))
    draws = [torch.from_numpy(z[f'draw_{i}']) for i in range(int(z['n_draws']))]
    out, hist = gd.sample(batch, return_history=True, noise=draws)
    hist = torch.stack(hist).cpu().numpy()
    assert hist.shape == z['history'].shape
    scale = np.abs(z['history']).reshape(T + 1, -1).max(1).clip(1.0)[:, None, None]
    err = float(np.max(np.abs(hist - z['history']) / scale))
    assert err < 2e-4, (name, err)
    mk = batch.mask.numpy().astype(bool)
    gt = batch.x.numpy()[:, dims[-1][1]:dims[-1][2]]
    assert all(np.array_equal(h[mk], gt[mk]) for h in hist)


def test_metropolis_samplers_need_the_energy_form():
    den = ConstraintDiffuser(dims=synthetic.DIMS['qualitative'], input_mode='qualitative', device='cuda', verbose=False)
    with pytest.raises(ValueError):
        GaussianDiffusion(den, timesteps=8, EBM='MALA')
