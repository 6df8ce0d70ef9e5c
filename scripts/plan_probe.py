"""Developer probe: where does the host-side time of one end-to-end sample() go (plan drop / build / sample)?"""
import sys
import time

import torch

sys.path.insert(0, '.')
from diffusion_ccsp_b200 import scenes, synthetic
from diffusion_ccsp_b200.ddpm import GaussianDiffusion
from diffusion_ccsp_b200.denoise_fn import ConstraintDiffuser

mode, dims = 'qualitative', synthetic.DIMS['qualitative']
sd = synthetic.make_trained_state_dict()
batch = scenes.qualitative_batch(1024, 8)
pinned = scenes.SceneBatch(batch.x.pin_memory(), batch.edge_index.pin_memory(), batch.edge_attr.pin_memory(), batch.mask.pin_memory())
den = ConstraintDiffuser(dims=dims, input_mode=mode, device='cuda:0', verbose=False, math='bf16x3')
T = int(sys.argv[1]) if len(sys.argv) > 1 else 300
gd = GaussianDiffusion(den, timesteps=T, EBM='ULA', samples_per_step=10).eval()
gd.load_state_dict(sd, strict=False)
for name, b in (('unpinned', batch), ('pinned', pinned), ('unpinned', batch), ('pinned', pinned)):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); den.drop_plans(); torch.cuda.synchronize()
    t1 = time.perf_counter(); den.plan_for(b); torch.cuda.synchronize()
    t2 = time.perf_counter(); gd.sample(b, seed=1)
    t3 = time.perf_counter()
    print(f'{name:9s} drop {1e3*(t1-t0):7.1f} ms | plan {1e3*(t2-t1):7.1f} ms | sample(T={T}) {1e3*(t3-t2):8.1f} ms')
