"""Generate golden vectors by running the UNMODIFIED reference sampler on seeded inputs.

Build-container only (needs /root/reference; see oracle/ref_shim.py).  For every case it stores the
inputs (scene batch arrays, seeds for weights/noise) and the reference's outputs in
tests/golden/<case>.npz.  Weights come from diffusion_ccsp_b200.synthetic.make_state_dict(seed)
(loaded into the reference with load_state_dict), noise from synthetic.make_noise(seed) served to the
reference by patching torch.randn (ddpm.py:121-122, 273, 292).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_shim import injected_randn, load_reference  # noqa: E402
from diffusion_ccsp_b200 import scenes, synthetic  # noqa: E402

MODES = {
    # case-name: (input_mode, triangular, scene factory)
    'qualitative': ('qualitative', False, lambda: scenes.qualitative_batch(8, 4)),
    'boxes': ('diffuse_pairwise', False, lambda: scenes.make_batch('boxes', 3, 12, seed=3)),
    'triangles': ('diffuse_pairwise', True, lambda: scenes.make_batch('triangles', 3, 10, seed=4)),
    'robot_box': ('robot_box', False, lambda: scenes.make_batch('robot_box', 3, 6, seed=5)),
    # the 3-type 'stability_flat' vocabulary (denoise_fn.py:18, 209-210; rows like qualitative: geom 2 + pose 4)
    'stability': ('stability_flat', False, lambda: scenes.collate([scenes.random_typed_scene(np.random.default_rng(60 + i), 5, 3, 14, 6)
                                                                   for i in range(3)])),
}
ONLY = [a for a in sys.argv[1:] if not a.startswith('--')]      # e.g. `make_golden.py stability`: only these MODES cases


def build_reference(input_mode, dims, T, EBM, K, weight_seed):
    dfn, ddpm = load_reference()
    m = dfn.ConstraintDiffuser(dims=dims, hidden_dim=256, input_mode=input_mode, EBM=EBM,
                               device='cpu', verbose=False)
    gd = ddpm.GaussianDiffusion(m, timesteps=T, EBM=EBM, samples_per_step=K,
                                step_sizes='2*self.betas').eval()
    sd = synthetic.make_state_dict(dims, input_mode, seed=weight_seed)
    missing, unexpected = gd.load_state_dict(sd, strict=False)
    assert not unexpected and all(not k.startswith('denoise_fn.') for k in missing), (missing, unexpected)
    return m, gd


def batch_arrays(b):
    return dict(x=b.x.numpy(), edge_index=b.edge_index.numpy().astype(np.int32),
                edge_attr=b.edge_attr.numpy().astype(np.int8), mask=b.mask.numpy())


def save(name, **arrs):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrs)
    print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB')


def ula_plus_case():
    """EBM='ULA+' (ddpm.py:297-299): 4/8/12/16 ULA steps per quarter of the schedule, low t -> high t."""
    b = scenes.qualitative_batch(4, 3)
    dims = synthetic.DIMS['qualitative']
    T = 8
    m, gd = build_reference('qualitative', dims, T, 'ULA+', 10, weight_seed=5)
    draws = 1 + T + 2 * (4 + 8 + 12 + 16)
    rng = np.random.default_rng(99)
    noise = torch.from_numpy(rng.standard_normal((draws, b.num_nodes, 4), dtype=np.float32))
    with injected_randn(noise) as inj:
        out, hist = gd.sample(b, return_history=True)
    assert inj.calls == draws, (inj.calls, draws)
    hist = torch.stack([h.detach() for h in hist]).numpy()
    save('ulaplus_qualitative_T8', out=out.detach().numpy(), history=hist, T=T, weight_seed=5, noise_seed=99,
         input_mode='qualitative', triangular=False, **batch_arrays(b))


def mode_cases(modes):
    # ---- single denoiser evaluation per mode (denoise_fn.py:453-537) ----------------------------
    for case, (mode, tri, factory) in modes.items():
        dims = synthetic.dims_for(mode, tri)
        b = factory()
        m, gd = build_reference(mode, dims, 100, 'ULA', 10, weight_seed=11)
        rng = np.random.default_rng(77)
        poses = rng.standard_normal((b.num_nodes, dims[-1][0])).astype(np.float32)
        outs = []
        tvals = np.array([0, 37, 99])
        for t in tvals:
            with torch.no_grad():
                o = m(torch.from_numpy(poses.copy()), b, torch.tensor([int(t)]), eval=True)
            outs.append(o.detach().numpy())
        save(f'forward_{case}', poses_in=poses, t=tvals, out=np.stack(outs), weight_seed=11,
             triangular=tri, input_mode=mode, **batch_arrays(b))

    # ---- short trajectories for the other modes + plain DDPM + other K --------------------------
    for case, (mode, tri, factory) in modes.items():
        dims = synthetic.dims_for(mode, tri)
        b = factory()
        for (T, EBM, K) in ((3, 'ULA', 10), (6, False, 0), (4, 'ULA', 3)):
            m, gd = build_reference(mode, dims, T, EBM, K if K else 10, weight_seed=21)
            noise = synthetic.make_noise(T, K, b.num_nodes, dims[-1][0], seed=456)
            with injected_randn(noise) as inj:
                out, hist = gd.sample(b, return_history=True)
            assert inj.calls == 1 + T * (1 + K), (inj.calls, T, K)
            hist = torch.stack([h.detach() for h in hist]).numpy()
            tag = f'K{K}' if EBM else 'ddpm'
            save(f'traj_{case}_T{T}_{tag}', out=out.detach().numpy(), history=hist, T=T, K=K,
                 EBM=str(EBM), weight_seed=21, noise_seed=456, input_mode=mode, triangular=tri,
                 **batch_arrays(b))


def main():
    torch.set_num_threads(8)
    dfn, ddpm = load_reference()
    if '--ulaplus-only' in sys.argv:
        return ula_plus_case()
    if ONLY:
        return mode_cases({k: v for k, v in MODES.items() if k in ONLY})

    # ---- schedule tables (ddpm.py:184-226) --------------------------------------------------
    for T in (100, 1000):
        m, gd = build_reference('qualitative', synthetic.DIMS['qualitative'], T, 'ULA', 10, 0)
        tabs = {k: v.numpy() for k, v in gd.state_dict().items() if not k.startswith('denoise_fn.')}
        tabs['_sqrt_recipm1_alphas_cumprod_custom'] = gd._sqrt_recipm1_alphas_cumprod_custom.numpy()
        tabs['step_sizes'] = gd.step_sizes.numpy()
        save(f'schedule_T{T}', **tabs)

    # ---- time embedding (denoise_fn.py:43-50, 259-264) ----------------------------------------
    m, gd = build_reference('qualitative', synthetic.DIMS['qualitative'], 1000, 'ULA', 10, 0)
    ts = np.array([0, 1, 2, 7, 50, 99, 100, 500, 998, 999])
    with torch.no_grad():
        te = m.time_mlp(torch.tensor(ts, dtype=torch.long)).numpy()
        pe = m.time_mlp[0](torch.tensor(ts, dtype=torch.long)).numpy()
    save('time_embedding', t=ts, time_mlp=te, pos_emb=pe, weight_seed=0)

    mode_cases(MODES)

    # ---- trajectories: config 1 (qualitative N=4, batch 8, ULA K=10) --------------------------
    b = scenes.qualitative_batch(8, 4)
    dims = synthetic.DIMS['qualitative']
    for T in (1, 5, 20, 100):
        m, gd = build_reference('qualitative', dims, T, 'ULA', 10, weight_seed=0)
        noise = synthetic.make_noise(T, 10, b.num_nodes, 4, seed=123)
        with injected_randn(noise) as inj:
            out, hist = gd.sample(b, return_history=True)
        assert inj.calls == 1 + T * 11 and len(hist) == T + 1
        hist = torch.stack([h.detach() for h in hist]).numpy()
        save(f'traj_qualitative_n4_T{T}', out=out.detach().numpy(), history=hist, T=T, K=10,
             weight_seed=0, noise_seed=123, input_mode='qualitative', triangular=False, **batch_arrays(b))

    ula_plus_case()


if __name__ == '__main__':
    main()
