"""Golden vectors for the energy-form variant from the UNMODIFIED reference:
ConstraintDiffuser(energy_wrapper=True) wrapped in ComposedEBMDenoiseFn (networks/denoise_fn.py:57-83, 518-521, 539-548) —
one (energy, gradient) evaluation per input mode, and short trajectories with EBM = 'ULA' / 'MALA' / 'HMC'
(networks/ddpm.py:955-966, 1000-1033, 1036-1126) under injected draws (torch.randn, torch.randn_like and torch.rand are served
from one recorded sequence).

    python tests/golden/make_energy_golden.py
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_shim import load_reference  # noqa: E402
from diffusion_ccsp_b200 import scenes, synthetic  # noqa: E402
from tests.golden.make_golden import MODES, batch_arrays, save  # noqa: E402


class recorded_draws:
    """patch torch.randn / randn_like / rand: every call draws from a seeded numpy stream and is recorded in order"""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.draws = []

    def __enter__(self):
        self._saved = (torch.randn, torch.randn_like, torch.rand)
        me = self

        def shape_of(args):
            return tuple(args[0]) if len(args) == 1 and not isinstance(args[0], int) else tuple(args)

        def randn(*a, **k):
            z = me.rng.standard_normal(shape_of(a)).astype(np.float32)
            me.draws.append(z)
            return torch.from_numpy(z.copy())

        def randn_like(x, **k):
            return randn(*x.shape)

        def rand(*a, **k):
            z = me.rng.random(shape_of(a)).astype(np.float32)
            me.draws.append(z)
            return torch.from_numpy(z.copy())

        torch.randn, torch.randn_like, torch.rand = randn, randn_like, rand
        return self

    def __exit__(self, *exc):
        torch.randn, torch.randn_like, torch.rand = self._saved
        torch.set_grad_enabled(True)
        return False


def build(mode, dims, T, EBM, K, weight_seed, step_sizes='2*self.betas'):
    dfn, ddpm = load_reference()
    m = dfn.ConstraintDiffuser(dims=dims, hidden_dim=256, input_mode=mode, EBM=EBM, energy_wrapper=True, device='cpu', verbose=False)
    w = dfn.ComposedEBMDenoiseFn(m, 1)
    gd = ddpm.GaussianDiffusion(w, timesteps=T, EBM=EBM, samples_per_step=K, step_sizes=step_sizes).eval()
    sd = {k.replace('denoise_fn.', 'denoise_fn.model.'): v for k, v in synthetic.make_state_dict(dims, mode, seed=weight_seed).items()}
    missing, unexpected = gd.load_state_dict(sd, strict=False)
    assert not unexpected and not [k for k in missing if k.startswith('denoise_fn.')], (missing, unexpected)
    return m, w, gd


STEP_SIZES = '0.002*self.betas'      # the untrained energy has gradients ~ 2 deg x: the default 2*betas step diverges to NaN in a few steps


def main():
    torch.set_num_threads(8)
    # ---- one (energy, gradient) evaluation per input mode ---------------------------------------------------------
    for case, (mode, tri, factory) in MODES.items():
        dims = synthetic.dims_for(mode, tri)
        b = factory()
        m, w, gd = build(mode, dims, 100, 'ULA', 10, weight_seed=41)
        rng = np.random.default_rng(7)
        poses = (0.7 * rng.standard_normal((b.num_nodes, dims[-1][0]))).astype(np.float32)
        grads, energies = [], []
        tvals = np.array([0, 41, 99])
        for t in tvals:
            torch.set_grad_enabled(True)
            g, e = m(torch.from_numpy(poses.copy()), b, torch.tensor([int(t)]), eval=True, tag='EBM')
            grads.append(g.detach().numpy()); energies.append(float(e.detach()))
        save(f'energy_{case}', poses_in=poses, t=tvals, grad=np.stack(grads), energy=np.array(energies, np.float64), weight_seed=41,
             triangular=tri, input_mode=mode, **batch_arrays(b))

    # ---- trajectories ---------------------------------------------------------------------------------------------------
    b = scenes.qualitative_batch(4, 3)
    dims = synthetic.DIMS['qualitative']
    for EBM, K, T in (('ULA', 3, 12), ('MALA', 3, 12), ('HMC', 4, 12)):
        m, w, gd = build('qualitative', dims, T, EBM, K, weight_seed=43, step_sizes=STEP_SIZES)
        # scale the decoder down so that the untrained energy landscape keeps the chain at O(1) (acceptance neither 0 nor 1)
        with torch.no_grad():
            m.pose_decoder[2].weight.mul_(0.05); m.pose_decoder[2].bias.mul_(0.05)
        with recorded_draws(500 + len(EBM)) as rec, contextlib.redirect_stdout(io.StringIO()):
            out, hist = gd.sample(b, return_history=True)
        hist = torch.stack([h.detach() for h in hist]).numpy()
        arrs = {f'draw_{i}': z for i, z in enumerate(rec.draws)}
        save(f'ebm_{EBM.lower()}_qualitative_T{T}', out=out.detach().numpy(), history=hist, T=T, K=K, EBM=EBM, weight_seed=43,
             decoder_scale=0.05, step_sizes=STEP_SIZES, n_draws=len(rec.draws), input_mode='qualitative', triangular=False, **arrs, **batch_arrays(b))
        print(f'   {EBM}: {len(rec.draws)} draws, max|x| {np.abs(hist).max():.3f}')


if __name__ == '__main__':
    main()
