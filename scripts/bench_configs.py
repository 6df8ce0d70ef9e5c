"""Per-evaluation cost of the sampling loop on the other BASELINE.json configs at their per-GPU batch sizes
(bench.py measures configs[1] only, by contract).  Short runs (T = 40), device-timed, extrapolated to T = 1000."""
import json
import sys
import time

import torch

sys.path.insert(0, '.')
from diffusion_ccsp_b200 import scenes, synthetic
from diffusion_ccsp_b200.ddpm import GaussianDiffusion
from diffusion_ccsp_b200.denoise_fn import ConstraintDiffuser

CASES = [('config 1: qualitative N=4, batch 8', 'qualitative', 'qualitative', False, 4, 8),
         ('config 2: qualitative N=8, batch 1024', 'qualitative', 'qualitative', False, 8, 1024),
         ('config 3: boxes N=12, batch 4096', 'boxes', 'diffuse_pairwise', False, 12, 4096),
         ('config 4: triangles N=10, batch 1024 (8192 / 8 GPUs)', 'triangles', 'diffuse_pairwise', True, 10, 1024),
         ('config 5: robot boxes N=6, batch 256 (2048 / 8 GPUs)', 'robot_box', 'robot_box', False, 6, 256),
         ('config 2 strong-scaling shard at 8 GPUs: qualitative N=8, batch 128', 'qualitative', 'qualitative', False, 8, 128)]
T, K = 40, 10
for name, kind, mode, tri, n_obj, B in CASES:
    dims = synthetic.dims_for(mode, tri)
    sd = synthetic.make_state_dict(dims, mode, seed=0)
    batch = scenes.qualitative_batch(B, n_obj) if kind == 'qualitative' else scenes.make_batch(kind, B, n_obj, seed=0)
    den = ConstraintDiffuser(dims=dims, input_mode=mode, device='cuda:0', verbose=False, math='bf16x3')
    gd = GaussianDiffusion(den, timesteps=T, EBM='ULA', samples_per_step=K).eval()
    gd.load_state_dict(sd, strict=False)
    gd.sample(batch, seed=1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(3):
        gd.sample(batch, seed=2 + i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    per_eval = ms / (T * (1 + K))
    print(json.dumps(dict(case=name, nodes=batch.num_nodes, edges=batch.num_edges, ms_per_eval=round(per_eval, 4),
                          scenes_per_s_at_T1000=round(B / (per_eval * 11000 / 1e3), 1))))
