"""Golden vectors for the training step from the UNMODIFIED reference's autograd.

For every input mode: loss = GaussianDiffusion.p_losses(batch, t, noise=..., debug=False, tag='EBM') (networks/ddpm.py:363-385)
with seeded weights / noise, then loss.backward().  Stored per case: the loss, the denoiser output, and for every parameter
the gradient's max-abs, its float64 sum of squares and 256 seeded sample entries (full tensors would be 36 MB per case).

    python tests/golden/make_train_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_shim import load_reference  # noqa: E402
from diffusion_ccsp_b200 import scenes, synthetic  # noqa: E402
from tests.golden.make_golden import MODES, batch_arrays, build_reference, save  # noqa: E402

N_SAMPLES = 256


def reference_loss_and_grads(mode, dims, batch, T, t, noise, weight_seed, loss_type='l2', sd=None):
    """returns (loss, recon [n,P], {name: grad ndarray or None})"""
    dfn, ddpm = load_reference()
    m = dfn.ConstraintDiffuser(dims=dims, hidden_dim=256, input_mode=mode, EBM='ULA', device='cpu', verbose=False)
    gd = ddpm.GaussianDiffusion(m, timesteps=T, loss_type=loss_type, EBM='ULA', samples_per_step=10)
    sd = sd if sd is not None else synthetic.make_state_dict(dims, mode, seed=weight_seed)
    missing, unexpected = gd.load_state_dict(sd, strict=False)
    assert not unexpected
    gd.train()
    torch.set_grad_enabled(True)
    captured = {}
    orig = m.forward

    def spy(*a, **k):
        out = orig(*a, **k)
        captured['recon'] = out.detach().clone()
        return out

    m.forward = spy
    loss = gd.p_losses(batch, torch.tensor([t]), noise=torch.from_numpy(noise.copy()), debug=False, tag='EBM')
    loss.backward()
    grads = {k: (p.grad.detach().numpy().copy() if p.grad is not None else None) for k, p in m.named_parameters()}
    return float(loss.detach()), captured['recon'].numpy(), grads


def sample_index(name, numel):
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    return rng.choice(numel, size=min(numel, N_SAMPLES), replace=False)


def main():
    torch.set_num_threads(8)
    for case, (mode, tri, factory) in MODES.items():
        dims = synthetic.dims_for(mode, tri)
        b = factory()
        P = dims[-1][0]
        for (t, loss_type) in ((37, 'l2'), (3, 'l1')):
            rng = np.random.default_rng(1000 + t)
            noise = rng.standard_normal((b.num_nodes, P)).astype(np.float32)
            noise[b.mask.numpy().astype(bool)] = 0                                       # conditional_noise (ddpm.py:114-117)
            loss, recon, grads = reference_loss_and_grads(mode, dims, b, 100, t, noise, weight_seed=31, loss_type=loss_type)
            arrs = dict(loss=np.float64(loss), recon=recon, noise=noise, t=t, T=100, weight_seed=31, input_mode=mode,
                        triangular=tri, loss_type=loss_type, **batch_arrays(b))
            for k, g in grads.items():
                if g is None:
                    arrs[f'none:{k}'] = np.int8(1)
                    continue
                idx = sample_index(k, g.size)
                arrs[f'idx:{k}'] = idx.astype(np.int64)
                arrs[f'val:{k}'] = g.reshape(-1)[idx]
                arrs[f'max:{k}'] = np.float64(np.abs(g).max())
                arrs[f'ssq:{k}'] = np.float64((g.astype(np.float64) ** 2).sum())
            save(f'train_{case}_{loss_type}', **arrs)


if __name__ == '__main__':
    main()
