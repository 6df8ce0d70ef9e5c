#!/usr/bin/env python
"""Small run of the round-2 kernels (solved-checker, training step, energy gradient) for compute-sanitizer:
    compute-sanitizer --tool memcheck python scripts/sanitize_probe.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusion_ccsp_b200 import scenes, synthetic
from diffusion_ccsp_b200.checker import SolvedChecker
from diffusion_ccsp_b200.ddpm import GaussianDiffusion
from diffusion_ccsp_b200.denoise_fn import ConstraintDiffuser
from diffusion_ccsp_b200.ebm import ComposedEBMDenoiseFn

dims = synthetic.DIMS['qualitative']
b = scenes.collate([scenes.qualitative_batch(6, 4), scenes.qualitative_batch(3, 8), scenes.qualitative_batch(2, 3)])
ck = SolvedChecker(b, dims, 'qualitative', 'cuda')
s, c = ck(b.x[:, 2:6], return_counts=True)
print('checker', int(s.sum()), 'of', s.numel())
m = ConstraintDiffuser(dims=dims, input_mode='qualitative', device='cuda', verbose=False)
gd = GaussianDiffusion(m, timesteps=50, EBM='ULA').train()
gd.load_state_dict(synthetic.make_state_dict(dims, 'qualitative', seed=1), strict=False)
m.to('cuda')
loss = gd.p_losses(b, 17, debug=False)
loss.backward()
print('train loss', float(loss.detach()), 'grad norm', float(sum(p.grad.pow(2).sum() for p in m.parameters() if p.grad is not None).sqrt()))
rb = scenes.make_batch('robot_box', 3, 5, seed=2)
rd = synthetic.dims_for('robot_box')
rm = ConstraintDiffuser(dims=rd, input_mode='robot_box', device='cuda', verbose=False)
rgd = GaussianDiffusion(rm, timesteps=50, EBM='ULA').train()
rgd.load_state_dict(synthetic.make_state_dict(rd, 'robot_box', seed=1), strict=False)
rm.to('cuda')
rgd.p_losses(rb, 3, debug=False).backward()
em = ConstraintDiffuser(dims=dims, input_mode='qualitative', energy_wrapper=True, device='cuda', verbose=False)
em.load_state_dict({k[len('denoise_fn.'):]: v for k, v in synthetic.make_state_dict(dims, 'qualitative', seed=1).items()})
em.to('cuda')
g, e = em(torch.randn(b.num_nodes, 4), b, torch.tensor([5]), tag='EBM')
torch.cuda.synchronize()
print('energy', float(e), 'ok')
# the sampling kernels themselves (node + fused edge kernel, BF16x3 and FP32 modes), ragged batch, a few evaluations
for math in ('bf16x3', 'fp32'):
    sm = ConstraintDiffuser(dims=dims, input_mode='qualitative', device='cuda', verbose=False, math=math)
    sgd = GaussianDiffusion(sm, timesteps=3, EBM='ULA', samples_per_step=2).eval()
    sgd.load_state_dict(synthetic.make_state_dict(dims, 'qualitative', seed=1), strict=False)
    out = sgd.sample(b, seed=7)
    torch.cuda.synchronize()
    print('sample', math, 'finite', bool(torch.isfinite(out).all()))
