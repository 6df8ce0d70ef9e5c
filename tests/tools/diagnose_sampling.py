#!/usr/bin/env python
"""TEST TOOLING (may load the staged reference through oracle/ref_shim.py).  Does a checkpoint sample stably?  Our CUDA sampler (bf16x3 and fp32) and, when staged, the unmodified reference on the CPU, on
the same few scenes with the same injected noise: max|x| along the trajectory and the first timestep that goes non-finite."""
import argparse, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from diffusion_ccsp_b200 import scenes, synthetic
from diffusion_ccsp_b200.ddpm import GaussianDiffusion
from diffusion_ccsp_b200.denoise_fn import ConstraintDiffuser

ap = argparse.ArgumentParser()
ap.add_argument('--ckpt', required=True)
ap.add_argument('--scenes', type=int, default=4)
ap.add_argument('--timesteps', type=int, default=1000)
ap.add_argument('--reference', action='store_true')
a = ap.parse_args()
dims = synthetic.DIMS['qualitative']
sd = {k: v.float() for k, v in torch.load(a.ckpt, map_location='cpu').items()}
batch = scenes.qualitative_batch(a.scenes, 8)
T, K = a.timesteps, 10
noise = synthetic.make_noise(T, K, batch.num_nodes, 4, seed=5)


def summarize(tag, hist):
    h = torch.stack([x.detach().float().cpu() for x in hist])                      # [T+1, n, P]
    mx = h.abs().amax(dim=(1, 2))
    bad = (~torch.isfinite(h)).any(dim=(1, 2)).nonzero()
    first_bad = int(bad[0]) if bad.numel() else None
    pts = [0, 1, 2, 5, 10, 20, 50, 100, 200, 400, 600, 800, 900, 950, 990, T]
    print(f'[{tag}] first non-finite history index: {first_bad};  max|x| at steps', {p: float(mx[min(p, T)]) for p in pts}, flush=True)
    return h


outs = {}
for math in ('bf16x3', 'fp32'):
    den = ConstraintDiffuser(dims=dims, input_mode='qualitative', device='cuda', verbose=False, math=math)
    gd = GaussianDiffusion(den, timesteps=T, EBM='ULA', samples_per_step=K).eval()
    gd.load_state_dict(sd, strict=False)
    out, hist = gd.sample(batch, return_history=True, noise=noise)
    outs[math] = summarize(math, hist)
if a.reference:
    from oracle import ref_shim
    dfn, ddpm = ref_shim.load_reference()
    m = dfn.ConstraintDiffuser(dims=dims, hidden_dim=256, input_mode='qualitative', EBM='ULA', device='cpu', verbose=False)
    rgd = ddpm.GaussianDiffusion(m, timesteps=T, EBM='ULA', samples_per_step=K).eval()
    rgd.load_state_dict(sd, strict=False)
    torch.set_num_threads(os.cpu_count())
    t0 = time.time()
    with ref_shim.injected_randn(noise):
        out, hist = rgd.sample(batch, return_history=True)
    print(f'reference: {time.time() - t0:.0f} s')
    ref = summarize('reference', hist)
    for math, h in outs.items():
        fin = torch.isfinite(ref).all(dim=(1, 2)) & torch.isfinite(h).all(dim=(1, 2))
        d = ((h - ref).abs().amax(dim=(1, 2)) / ref.abs().amax(dim=(1, 2)).clamp_min(1.0))[fin]
        print(f'[{math} vs reference] max rel diff over the finite prefix ({int(fin.sum())} states): {float(d.max()):.3e}')
