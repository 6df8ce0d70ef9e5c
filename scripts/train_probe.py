#!/usr/bin/env python
"""Where does a training step spend its time?  (host loader / graph build / kernels / optimiser)"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusion_ccsp_b200 import scenes, synthetic, train
from diffusion_ccsp_b200.ddpm import GaussianDiffusion
from diffusion_ccsp_b200.denoise_fn import ConstraintDiffuser

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dims = synthetic.DIMS['qualitative']
pool = scenes.qualitative_batch(1024, 8)
m = ConstraintDiffuser(dims=dims, input_mode='qualitative', device='cuda', verbose=False)
gd = GaussianDiffusion(m, timesteps=1000, EBM='ULA').train()
gd.load_state_dict(synthetic.make_state_dict(dims, 'qualitative', seed=0), strict=False)
m.to('cuda')
opt = train.Adam(gd.parameters(), lr=5e-4, on_step=m.mark_weights_dirty)
ld = scenes.SceneLoader(pool, B, shuffle=True)
T = dict(load=0.0, graph=0.0, fwdbwd=0.0, backward=0.0, opt=0.0, kernels=0.0)
it = iter(ld)
for s in range(steps + 3):
    if s == 3:
        for k in T: T[k] = 0.0
    torch.cuda.synchronize(); t0 = time.perf_counter()
    try: data = next(it)
    except StopIteration:
        it = iter(ld); data = next(it)
    t1 = time.perf_counter()
    g = m.train_graph_for(data)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = gd.p_losses(data, int(np.random.randint(0, 1000)), debug=False)
    e1.record()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    loss.backward()
    torch.cuda.synchronize(); t4 = time.perf_counter()
    opt.step(); opt.zero_grad()
    torch.cuda.synchronize(); t5 = time.perf_counter()
    T['load'] += t1 - t0; T['graph'] += t2 - t1; T['fwdbwd'] += t3 - t2; T['backward'] += t4 - t3; T['opt'] += t5 - t4
    T['kernels'] += e0.elapsed_time(e1) * 1e-3
print(f'batch {B} scenes ({data.num_edges} edges): ' + '  '.join(f'{k} {v / steps * 1e3:.2f} ms' for k, v in T.items()),
      f' total {sum(v for k, v in T.items() if k != "kernels") / steps * 1e3:.2f} ms/step')
