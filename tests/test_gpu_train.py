"""GPU: the training step (ccsp_train_step / ccsp_adam_step through the C ABI, via GaussianDiffusion.p_losses) against
  (1) golden vectors from the UNMODIFIED reference's autograd (tests/golden/train_*.npz, made by make_train_golden.py),
  (2) the reference itself run live on the host CPU when its sources are staged (oracle/_ref travels to the GPU box):
      every element of every gradient, and a 60-step loss curve with Adam.

Stated FP32 tolerance: loss 2e-6 relative; denoiser output 5e-6 (max|d| / max(1, max|ref|)); gradients 2e-5 of the tensor's
max|grad| (FP32 FMA, fixed summation order; differs from the reference only in summation order).
"""
import numpy as np
import pytest
import torch

from diffusion_ccsp_b200 import scenes, synthetic, train
from diffusion_ccsp_b200.ddpm import GaussianDiffusion
from diffusion_ccsp_b200.denoise_fn import ConstraintDiffuser
from oracle import ref_shim
from tests.util import case_model, golden_names, load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL_LOSS, TOL_OUT, TOL_GRAD = 2e-6, 5e-6, 2e-5


def record(test, what, err):
    import json, os
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    if os.path.isdir(d):
        with open(os.path.join(d, 'train_parity_errors.jsonl'), 'a') as f:
            f.write(json.dumps(dict(test=test, what=what, err=err)) + '\n')


def build(mode, dims, sd, T=100, loss_type='l2'):
    m = ConstraintDiffuser(dims=dims, input_mode=mode, device='cuda', verbose=False, math='bf16x3')
    gd = GaussianDiffusion(m, timesteps=T, loss_type=loss_type, EBM='ULA', samples_per_step=10)
    missing, unexpected = gd.load_state_dict(sd, strict=False)
    assert not unexpected
    gd.denoise_fn.to('cuda')
    return m, gd.train()


@pytest.mark.parametrize('name', golden_names('train_'))
def test_loss_and_gradients_vs_reference_golden(name):
    z, batch = load_golden(name)
    mode, dims, sd = case_model(z)
    m, gd = build(mode, dims, sd, T=int(z['T']), loss_type=str(z['loss_type']))
    recon = torch.empty((batch.num_nodes, dims[-1][0]), device='cuda')
    loss = gd.p_losses(batch, torch.tensor([int(z['t'])]), noise=torch.from_numpy(z['noise']), debug=False, tag='EBM', recon=recon)
    loss.backward()
    e_loss = abs(float(loss) - float(z['loss'])) / abs(float(z['loss']))
    e_out = rel_err(recon.cpu().numpy(), z['recon'])
    record(name, 'loss', e_loss); record(name, 'recon', e_out)
    assert e_loss < TOL_LOSS, (float(loss), float(z['loss']))
    assert e_out < TOL_OUT, e_out
    worst = 0.0
    for k, p in m.named_parameters():
        if f'none:{k}' in z:
            assert p.grad is None, f'{k}: the reference leaves this gradient None (type without edges)'
            continue
        g = p.grad.detach().cpu().numpy().reshape(-1)
        gmax = float(z[f'max:{k}'])
        err = float(np.max(np.abs(g[z[f'idx:{k}']] - z[f'val:{k}']))) / max(gmax, 1e-30)
        ssq = float((g.astype(np.float64) ** 2).sum())
        worst = max(worst, err)
        assert err < TOL_GRAD, (k, err)
        assert abs(ssq - float(z[f'ssq:{k}'])) <= 1e-4 * float(z[f'ssq:{k}']) + 1e-30, (k, ssq, float(z[f'ssq:{k}']))
        assert abs(float(np.abs(g).max()) - gmax) <= 1e-4 * gmax + 1e-30, k
    record(name, 'grad_worst', worst)


def test_every_gradient_element_vs_live_reference():
    if not ref_shim.reference_available():
        pytest.skip('reference sources not staged')
    from tests.golden.make_train_golden import reference_loss_and_grads
    mode, dims = 'qualitative', synthetic.DIMS['qualitative']
    batch = scenes.collate([scenes.qualitative_batch(16, 4), scenes.qualitative_batch(16, 8), scenes.qualitative_batch(8, 3)])
    sd = synthetic.make_state_dict(dims, mode, seed=77)
    rng = np.random.default_rng(5)
    noise = rng.standard_normal((batch.num_nodes, 4)).astype(np.float32)
    noise[batch.mask.numpy().astype(bool)] = 0
    for t in (0, 61, 99):
        loss_ref, recon_ref, grads_ref = reference_loss_and_grads(mode, dims, batch, 100, t, noise, 77, sd=sd)
        m, gd = build(mode, dims, sd)
        loss = gd.p_losses(batch, torch.tensor([t]), noise=torch.from_numpy(noise), debug=False)
        loss.backward()
        assert abs(float(loss) - loss_ref) / abs(loss_ref) < TOL_LOSS
        worst = 0.0
        for k, p in m.named_parameters():
            ref = grads_ref[k]
            if ref is None:
                assert p.grad is None
                continue
            err = float(np.max(np.abs(p.grad.cpu().numpy() - ref))) / max(float(np.abs(ref).max()), 1e-30)
            worst = max(worst, err)
            assert err < TOL_GRAD, (t, k, err)
        record('live_full_grads', f't={t}', worst)


def test_step_is_bit_reproducible_and_accumulates():
    mode, dims = 'qualitative', synthetic.DIMS['qualitative']
    batch = scenes.qualitative_batch(32, 6)
    sd = synthetic.make_state_dict(dims, mode, seed=3)
    noise = synthetic.make_noise(0, 0, batch.num_nodes, 4, seed=9)[0]
    m, gd = build(mode, dims, sd)
    gd.p_losses(batch, 11, noise=noise, debug=False).backward()
    g1 = {k: p.grad.clone() for k, p in m.named_parameters()}
    m.zero_grad(set_to_none=True)
    l2 = gd.p_losses(batch, 11, noise=noise, debug=False)
    (l2 / 2).backward()
    (gd.p_losses(batch, 11, noise=noise, debug=False) / 2).backward()          # ddpm.py:534 gradient accumulation
    for k, p in m.named_parameters():
        assert torch.equal(p.grad, g1[k] * 0.5 + g1[k] * 0.5), k


def test_adam_matches_torch_adam():
    torch.manual_seed(0)
    p_ref = torch.nn.Parameter(torch.randn(5000, device='cuda'))
    p_own = torch.nn.Parameter(p_ref.detach().clone())
    o_ref = torch.optim.Adam([p_ref], lr=5e-4, foreach=False, fused=False)
    o_own = train.Adam([p_own], lr=5e-4)
    for i in range(50):
        g = torch.randn(5000, device='cuda') * (1.0 + i)
        p_ref.grad = g.clone(); p_own.grad = g.clone()
        o_ref.step(); o_own.step()
    assert float((p_ref - p_own).abs().max()) < 1e-6
    # a parameter without a gradient is skipped entirely (torch semantics)
    p_own.grad = None
    before = p_own.detach().clone()
    o_own.step()
    assert torch.equal(before, p_own.detach())


def test_loss_curve_vs_live_reference_with_adam():
    """60 Adam steps on the same batches / timesteps / noise: the loss curves agree"""
    if not ref_shim.reference_available():
        pytest.skip('reference sources not staged')
    dfn, ddpm = ref_shim.load_reference()
    mode, dims = 'qualitative', synthetic.DIMS['qualitative']
    pool = scenes.qualitative_batch(64, 4)
    sd = synthetic.make_state_dict(dims, mode, seed=13)
    ref_m = dfn.ConstraintDiffuser(dims=dims, hidden_dim=256, input_mode=mode, EBM='ULA', device='cpu', verbose=False)
    ref_gd = ddpm.GaussianDiffusion(ref_m, timesteps=100, EBM='ULA', samples_per_step=10)
    ref_gd.load_state_dict(sd, strict=False)
    ref_gd.train()
    torch.set_grad_enabled(True)
    ref_opt = torch.optim.Adam(ref_gd.parameters(), lr=5e-4)
    m, gd = build(mode, dims, sd)
    opt = train.Adam(gd.parameters(), lr=5e-4, on_step=m.mark_weights_dirty)
    rng = np.random.default_rng(21)
    ours, theirs = [], []
    for step in range(60):
        ids = rng.choice(64, 16, replace=False)
        batch = scenes.take_scenes(pool, ids)
        t = int(rng.integers(0, 100))
        noise = rng.standard_normal((batch.num_nodes, 4)).astype(np.float32)
        noise[batch.mask.numpy().astype(bool)] = 0
        lr = ref_gd.p_losses(batch, torch.tensor([t]), noise=torch.from_numpy(noise.copy()), debug=False, tag='EBM')
        ref_opt.zero_grad(); lr.backward(); ref_opt.step()
        lo = gd.p_losses(batch, t, noise=torch.from_numpy(noise), debug=False, tag='EBM')
        opt.zero_grad(); lo.backward(); opt.step()
        ours.append(float(lo)); theirs.append(float(lr))
    ours, theirs = np.array(ours), np.array(theirs)
    rel = np.abs(ours - theirs) / np.abs(theirs)
    record('loss_curve_60_steps', 'max_rel', float(rel.max()))
    assert theirs[-10:].mean() < 0.7 * theirs[:10].mean(), 'the run did not train'
    assert rel.max() < 2e-3, rel.max()
