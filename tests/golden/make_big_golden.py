"""Full-size goldens from the UNMODIFIED reference (build container, minutes of CPU):

  big_qualitative_n8_T1000   64 scenes x N = 8 x T = 1000 x ULA K = 10 with the checkpoint this repo trained (realistic regime)
  big_{boxes,triangles,robot_box}_T100   32 scenes at the configs' object counts, T = 100, K = 10, seeded-init weights (stress regime)

Only the final poses are stored.    python tests/golden/make_big_golden.py [case ...]
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
from tests.golden.make_golden import batch_arrays, save  # noqa: E402


def run(name, mode, tri, batch, sd, T, K, noise_seed, extra):
    dfn, ddpm = load_reference()
    dims = synthetic.dims_for(mode, tri)
    m = dfn.ConstraintDiffuser(dims=dims, hidden_dim=256, input_mode=mode, EBM='ULA', device='cpu', verbose=False)
    gd = ddpm.GaussianDiffusion(m, timesteps=T, EBM='ULA', samples_per_step=K).eval()
    missing, unexpected = gd.load_state_dict(sd, strict=False)
    assert not unexpected
    noise = synthetic.make_noise(T, K, batch.num_nodes, dims[-1][0], seed=noise_seed)
    t0 = time.time()
    # no history: the reference's history list keeps every timestep's ULA autograd graph alive (sample_step runs under
    # enable_grad, ddpm.py:955) — tens of GB at these sizes
    with injected_randn(noise) as inj:
        out = gd.sample(batch)
    assert inj.calls == 1 + T * (1 + K)
    print(f'{name}: {time.time() - t0:.0f} s, max|x| {np.abs(out.detach().numpy()).max():.3f}')
    save(name, out=out.detach().numpy(), T=T, K=K, noise_seed=noise_seed, input_mode=mode, triangular=tri, **extra,
         **batch_arrays(batch))


def main():
    torch.set_num_threads(os.cpu_count())
    want = set(sys.argv[1:])
    if not want or 'qualitative' in want:
        run('big_qualitative_n8_T1000', 'qualitative', False, scenes.qualitative_batch(64, 8), synthetic.load_trained_checkpoint(),
            1000, 10, 321, dict(weights='trained_checkpoint'))
    for case, mode, tri, n_obj, seed in (('boxes', 'diffuse_pairwise', False, 12, 13), ('triangles', 'diffuse_pairwise', True, 10, 14),
                                         ('robot_box', 'robot_box', False, 6, 15)):
        if want and case not in want:
            continue
        dims = synthetic.dims_for(mode, tri)
        run(f'big_{case}_T100', mode, tri, scenes.make_batch(case, 32, n_obj, seed=seed), synthetic.make_state_dict(dims, mode, seed=51),
            100, 10, 322, dict(weight_seed=51))


if __name__ == '__main__':
    main()
