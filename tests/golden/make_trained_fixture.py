"""Train a SMALL subset of the reference model's parameters with the reference's OWN loss so that sampling
stays in the realistic O(1) regime, and save only that subset as a fixture.

Random-init weights make the sampler diverge (|x| ~ 1e4 after T=100: a stress case that amplifies error);
the authors' checkpoints are not available offline.  Training everything would need a 35 MB checkpoint, so
only pose_encoder, pose_decoder and the mlps' biases (~70 k parameters) are trained; every other tensor stays
at synthetic.make_state_dict(seed) values and is regenerated from the seed.

Build-container only.   python tests/golden/make_trained_fixture.py
Output: diffusion_ccsp_b200/data/trained_small_qualitative.npz  +  goldens trained_traj_qualitative_T{100,1000}.npz
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_shim import injected_randn, load_reference  # noqa: E402
from diffusion_ccsp_b200 import scenes, synthetic  # noqa: E402

SEED, T_TRAIN, STEPS = 7, 100, int(os.environ.get('STEPS', 1500))


def trainable(name):
    return name.startswith(('denoise_fn.pose_encoder', 'denoise_fn.pose_decoder')) or (name.startswith('denoise_fn.mlps') and name.endswith('bias'))


def main():
    torch.manual_seed(0); np.random.seed(0)
    torch.set_num_threads(8)
    dfn, ddpm = load_reference()
    mode, dims = 'qualitative', synthetic.DIMS['qualitative']
    m = dfn.ConstraintDiffuser(dims=dims, hidden_dim=256, input_mode=mode, EBM='ULA', device='cpu', verbose=False)
    gd = ddpm.GaussianDiffusion(m, timesteps=T_TRAIN, EBM='ULA', samples_per_step=10)
    gd.load_state_dict(synthetic.make_state_dict(dims, mode, seed=SEED), strict=False)
    params = []
    for n, p in gd.named_parameters():
        p.requires_grad_(trainable(n))
        if trainable(n):
            params.append(p)
    print('trainable parameters:', sum(p.numel() for p in params))
    opt = torch.optim.Adam(params, lr=5e-4)
    pools = [scenes.qualitative_batch(64, 4), scenes.qualitative_batch(16, 3), scenes.qualitative_batch(16, 6),
             scenes.qualitative_batch(64, 8, seed=3)]
    gd.train()
    t0 = time.time()
    for step in range(STEPS):
        batch = pools[step % len(pools)]
        torch.set_grad_enabled(True)
        loss = gd(batch, debug=False, tag='EBM')
        opt.zero_grad(); loss.backward(); opt.step()
        if step % 100 == 0:
            print(f'step {step} loss {loss.item():.4f} ({time.time() - t0:.0f}s)', flush=True)
    gd.eval()
    sd = {k: v.detach().numpy().copy() for k, v in gd.state_dict().items() if trainable(k)}
    np.savez_compressed(os.path.join(os.path.dirname(os.path.dirname(HERE)), 'diffusion_ccsp_b200', 'data', 'trained_small_qualitative.npz'), weight_seed=SEED, **sd)
    print('saved', sum(v.size for v in sd.values()), 'floats')

    # goldens in the realistic regime: config 1 shape, T = 100 (and T = 1000 on 2 scenes)
    for T, nscenes in ((100, 8), (1000, 2)):
        gdT = ddpm.GaussianDiffusion(m, timesteps=T, EBM='ULA', samples_per_step=10).eval()
        b = scenes.qualitative_batch(nscenes, 4)
        noise = synthetic.make_noise(T, 10, b.num_nodes, 4, seed=321)
        t0 = time.time()
        with injected_randn(noise):
            out, hist = gdT.sample(b, return_history=True)
        out = out.detach().numpy()
        print(f'T={T}: max|x| = {np.abs(out).max():.3f}, frac in [-1.05,1.05] = {(np.abs(out) < 1.05).mean():.3f}, {time.time() - t0:.0f}s')
        hist = torch.stack([h.detach() for h in hist]).numpy()[:: max(1, T // 20)]
        np.savez_compressed(os.path.join(HERE, f'trained_traj_qualitative_T{T}.npz'), out=out, history_every=max(1, T // 20),
                            history=hist, T=T, K=10, noise_seed=321, n_scenes=nscenes, x=b.x.numpy(),
                            edge_index=b.edge_index.numpy().astype(np.int32), edge_attr=b.edge_attr.numpy().astype(np.int8),
                            mask=b.mask.numpy())


if __name__ == '__main__':
    main()
