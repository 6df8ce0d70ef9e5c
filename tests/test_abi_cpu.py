"""CPU: the C-ABI library builds, loads, and exports every symbol include/ccsp_b200.h declares
(no compute calls without a GPU), and fails loudly instead of falling back when CUDA is absent."""
import ctypes
import os
import re

import pytest
import torch

from diffusion_ccsp_b200 import _abi, build, synthetic
from diffusion_ccsp_b200.denoise_fn import ConstraintDiffuser

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib_path():
    return build.build_library()


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'ccsp_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ccsp_[a-z_0-9]+)\s*\(', src)))


def test_header_symbols_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/ccsp_b200.h but not exported'
    assert sorted(_abi.EXPORTS) == names
    lib.ccsp_abi_version.restype = ctypes.c_int
    assert lib.ccsp_abi_version() == 1


def test_library_is_sm100a_native(lib_path):
    import subprocess
    out = subprocess.run(['cuobjdump', '-lelf', lib_path], capture_output=True, text=True).stdout
    assert 'sm_100a' in out


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_no_cpu_fallback(lib_path):
    dims = synthetic.DIMS['qualitative']
    m = ConstraintDiffuser(dims=dims, input_mode='qualitative', device='cuda', verbose=False)
    from diffusion_ccsp_b200 import scenes
    b = scenes.qualitative_batch(2, 4)
    with pytest.raises(_abi.CcspError):
        m(torch.zeros(b.num_nodes, 4), b, torch.tensor([3]))


def test_state_dict_keys_match_reference_contract():
    """SURVEY.md §8b checkpoint key contract."""
    from diffusion_ccsp_b200.ddpm import GaussianDiffusion
    for mode, tri in (('qualitative', False), ('diffuse_pairwise', False), ('diffuse_pairwise', True), ('robot_box', False)):
        dims = synthetic.dims_for(mode, tri)
        m = ConstraintDiffuser(dims=dims, input_mode=mode, device='cuda', verbose=False)
        gd = GaussianDiffusion(m, timesteps=10, EBM='ULA')
        keys = set(gd.state_dict().keys())
        sd = synthetic.make_state_dict(dims, mode)
        assert set(sd.keys()) <= keys
        sched = {'betas', 'alphas_cumprod', 'alphas_cumprod_prev', 'sqrt_alphas_cumprod',
                 'sqrt_one_minus_alphas_cumprod', 'log_one_minus_alphas_cumprod', 'sqrt_recip_alphas_cumprod',
                 'sqrt_recipm1_alphas_cumprod', 'posterior_variance', 'posterior_log_variance_clipped',
                 'posterior_mean_coef1', 'posterior_mean_coef2'}
        assert keys == sched | set(sd.keys())
        for k, v in sd.items():
            assert gd.state_dict()[k].shape == v.shape, k
        gd.load_state_dict(sd, strict=False)


def test_plan_cache_never_returns_a_plan_of_other_content(monkeypatch):
    """the Trainer.evaluate pattern `batch = data.clone()` over different `data` recycles ids and
    storage addresses; the cache must never hand back a plan compiled from other content.  CPU test with a fake Plan."""
    import torch
    from diffusion_ccsp_b200 import _abi, scenes, synthetic
    from diffusion_ccsp_b200.denoise_fn import ConstraintDiffuser

    class FakePlan:
        def __init__(self, model, x, edge_index, edge_attr, mask, pose_begin, grasp_begin=0):
            self.content = (x.clone(), edge_index.clone(), edge_attr.clone(), mask.clone())
            self.closed = False

        def close(self):
            self.closed = True

    monkeypatch.setattr(_abi, 'Plan', FakePlan)
    den = ConstraintDiffuser(dims=synthetic.DIMS['qualitative'], input_mode='qualitative', device='cpu', verbose=False)
    monkeypatch.setattr(den, 'abi_model', lambda: object())
    pool = scenes.qualitative_batch(64, 4)
    hits = 0
    for it in range(200):
        data = pool.select_scenes(it % 60, it % 60 + 4)
        for _ in range(2):
            batch = data.clone()                       # freed at the end of the iteration -> addresses get recycled
            plan = den.plan_for(batch, verify_content=bool(it & 1))
            assert not plan.closed
            for a, b in zip(plan.content, (batch.x, batch.edge_index, batch.edge_attr, batch.mask)):
                assert torch.equal(a, b), 'stale plan returned for a different batch'
            again = den.plan_for(batch)
            hits += again is plan
    assert hits == 400                                  # the same live batch does hit the cache
    # an in-place edit through a numpy view leaves _version unchanged: the digest check of p_sample_loop catches it
    batch = pool.select_scenes(0, 4)
    p1 = den.plan_for(batch, verify_content=True)
    batch.x.numpy()[1, 2] += 0.25
    p2 = den.plan_for(batch, verify_content=True)
    assert p2 is not p1 and p1.closed and torch.equal(p2.content[0], batch.x)


def test_analysis_methods_match_the_reference():
    """`_get_constraint_inputs` / `_process_constraint` / `_add_constraints_outputs` (visualize_energy.py:401-455 uses them):
    composing them the way ConstraintDiffuser.forward does (denoise_fn.py:508-533) reproduces the reference's output."""
    import numpy as np
    import torch
    from diffusion_ccsp_b200 import scenes, synthetic
    from diffusion_ccsp_b200.denoise_fn import ConstraintDiffuser
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip('reference sources not available')
    dfn, _ = ref_shim.load_reference()
    for mode, batch in (('qualitative', scenes.qualitative_batch(3, 4)), ('robot_box', scenes.make_batch('robot_box', 2, 4, seed=1))):
        dims = synthetic.dims_for(mode)
        sd = {k[len('denoise_fn.'):]: v for k, v in synthetic.make_state_dict(dims, mode, seed=4).items()}
        ours = ConstraintDiffuser(dims=dims, input_mode=mode, device='cpu', verbose=False)
        ours.load_state_dict(sd)
        ref = dfn.ConstraintDiffuser(dims=dims, hidden_dim=256, input_mode=mode, device='cpu', verbose=False)
        ref.load_state_dict(sd)
        P = dims[-1][0]
        poses = torch.from_numpy(np.random.default_rng(0).standard_normal((batch.num_nodes, P)).astype(np.float32))
        t = torch.tensor([17])
        with torch.no_grad():
            want = ref(poses.clone(), batch, t, eval=True)
            emb = {'geoms_emb': ours.geom_encoder(batch.x[:, :dims[0][0]]), 'poses_emb': ours.pose_encoder(poses)}
            if 'robot' in mode:
                emb['grasp_emb'] = ours.grasp_encoder(batch.x[:, dims[1][1]:dims[1][2]])
            out, cnt = torch.zeros_like(poses), torch.zeros(batch.num_nodes)
            for i in range(len(ours.constraint_sets)):
                inp = ours._get_constraint_inputs(i, batch, t, emb, batch.edge_index.T)
                if inp['args'].shape[0] == 0:
                    continue
                o = ours._process_constraint(i, inp)
                ref_o = ref._process_constraint(i, ref._get_constraint_inputs(i, batch, t, emb, batch.edge_index.T))
                assert torch.equal(o, ref_o)
                assert torch.equal(ours._compute_energy(i, inp, poses, o), ref._compute_energy(i, inp, poses, o))
                out, cnt = ours._add_constraints_outputs(i, inp, o, out, cnt)
            out = out / torch.sqrt(cnt)[:, None]
            out[batch.mask.bool()] = batch.x[:, -P:][batch.mask.bool()]
        assert torch.allclose(out, want, rtol=0, atol=1e-6)


# ---- host logic of the pipelined-chains plan layout (runs without a device): where ccsp_plan_create cuts a batch -----------------
def _chain_cuts(lib_path, batch, want, num_types=13):
    import numpy as np
    lib = ctypes.CDLL(lib_path)
    lib.ccsp_debug_chain_cuts.restype = ctypes.c_int
    ei = batch.edge_index.to(torch.int64).contiguous()
    ea = batch.edge_attr.to(torch.float32).contiguous()
    out = np.zeros(want + 1, dtype=np.int64)
    k = lib.ccsp_debug_chain_cuts(ctypes.c_void_p(ei.data_ptr()), ctypes.c_void_p(ea.data_ptr()), ctypes.c_int64(batch.num_nodes),
                                  ctypes.c_int64(ei.shape[1]), ctypes.c_int32(num_types), ctypes.c_int32(want),
                                  ctypes.c_void_p(out.ctypes.data))
    return k, out[:k + 1].tolist()


@pytest.mark.parametrize('want', [2, 3, 4])
def test_chain_cuts_fall_between_scenes_and_balance_the_edges(lib_path, want):
    from diffusion_ccsp_b200 import scenes
    batch = scenes.qualitative_batch(40, 8)
    k, bounds = _chain_cuts(lib_path, batch, want)
    assert k == want and bounds[0] == 0 and bounds[-1] == batch.num_nodes and bounds == sorted(set(bounds))
    starts = set(batch.scene_node_ranges().tolist())
    assert all(b in starts for b in bounds)                       # only whole scenes
    i, j = batch.edge_index[0], batch.edge_index[1]
    per_chain = []
    for c in range(k):
        inside_i = (i >= bounds[c]) & (i < bounds[c + 1])
        inside_j = (j >= bounds[c]) & (j < bounds[c + 1])
        assert bool((inside_i == inside_j).all())                 # no edge crosses a cut
        per_chain.append(int(inside_i.sum()))
    assert sum(per_chain) == batch.num_edges
    assert max(per_chain) - min(per_chain) <= 2 * 90              # balanced to within about a scene (<= 90 edges at N = 8)


def test_chain_cuts_degenerate_cases(lib_path):
    from diffusion_ccsp_b200 import scenes
    one = scenes.qualitative_batch(1, 4)
    assert _chain_cuts(lib_path, one, 2) == (1, [0, one.num_nodes])         # a single scene cannot be cut
    two = scenes.qualitative_batch(2, 4)
    k, bounds = _chain_cuts(lib_path, two, 2)
    assert k == 2 and bounds[1] == int(two.scene_node_ranges()[1])
    assert _chain_cuts(lib_path, two, 4)[0] == 1                            # fewer legal cuts than asked for: stay whole
    assert _chain_cuts(lib_path, two, 1) == (1, [0, two.num_nodes])
