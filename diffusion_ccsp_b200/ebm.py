"""Energy-form variant of the sampler: ComposedEBMDenoiseFn (networks/denoise_fn.py:57-83), the energy branch
of ConstraintDiffuser.forward (:518-521, 539-548) and the annealed MCMC samplers of networks/ddpm.py:917-1128 (ULA on the energy
gradient, MALA, HMC / MUHA).

The heavy part — the energy E(x) = sum over edges |decoder(...) - x[arg]|^2 and its gradient dE/dx — is ONE C-ABI call
(`ccsp_energy_grad`, hand-written FP32 CUDA with the analytic input gradient; the reference uses torch.autograd.grad).  The
Metropolis bookkeeping around it (proposal, log-probabilities, accept / reject) is elementwise work on [n, P] tensors and stays in
host-driven torch ops on the device, step by step like the reference: these variants are only enabled by `-EBM MALA/HMC` or two
hard-coded run ids (train_utils.py:115-116, 333-334) and are not the headline path.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import _abi
from .train import Params, _declare, _pack, param_names


def energy_and_gradient(denoise_fn, batch, poses_in: torch.Tensor, t: int):
    """(dE/dx [n,P], E scalar) on the CUDA device — denoise_fn.py:518-521, 539-548, 373-375."""
    graph = denoise_fn.train_graph_for(batch)
    lib = _declare(_abi.load_library())
    lib.ccsp_energy_grad.argtypes = [C.c_void_p, C.POINTER(Params), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    dev = graph.device
    names = param_names(denoise_fn)
    sd = dict(denoise_fn.named_parameters())
    weights = {}
    for n in names:
        p = sd[n].detach()
        if not p.is_cuda:
            raise _abi.CcspError(f'parameter {n} is on {p.device}: move the model to the CUDA device first; the energy form has no CPU fallback')
        weights[n] = p
    w, keep = _pack(weights, graph.num_types)
    x = poses_in.detach().to(dev, torch.float32).contiguous()
    energy = torch.empty((), dtype=torch.float32, device=dev)
    grad = torch.empty_like(x)
    with torch.cuda.device(dev):
        _abi.check(lib.ccsp_energy_grad(graph._h, C.byref(w), int(t), x.data_ptr(), energy.data_ptr(), grad.data_ptr(),
                                        _abi.current_stream_ptr(dev)), 'ccsp_energy_grad')
    del keep
    return grad, energy


class ComposedEBMDenoiseFn(nn.Module):
    """denoise_fn.py:57-83: wrapper that exposes the energy gradient as the 'denoiser output'.  State-dict keys gain the
    `model.` prefix exactly like the reference (denoise_fn.model.*)."""

    def __init__(self, model, ebm_per_steps=1):
        super().__init__()
        self.model = model
        self.device = model.device
        self.dims = model.dims
        self.input_mode = model.input_mode
        self.ebm_per_steps = ebm_per_steps
        self.energy_wrapper = True

    def neg_logp_unnorm(self, poses_in, batch, t, **kwargs):
        kwargs['tag'] = 'EBM'
        _, energy = self.model.forward(poses_in, batch, t, **kwargs)
        return energy.sum()

    def forward(self, poses_in, batch, t, **kwargs):
        if isinstance(poses_in, np.ndarray):
            poses_in = torch.tensor(poses_in)
            t = torch.tensor(t)
        kwargs['tag'] = 'EBM'
        gradients, _ = self.model.forward(poses_in, batch, t, **kwargs)
        return gradients


# ---------------------------------------------------------------------------------------------------------------------
# noise sources: the reference calls torch.randn / torch.randn_like / torch.rand in a fixed order (ddpm.py:121-122, 273, 292,
# 1005, 1020, 1080, 1087, 1106); parity tests inject the same pre-drawn sequence on both sides
# ---------------------------------------------------------------------------------------------------------------------
class TorchNoise:
    def __init__(self, device, generator: Optional[torch.Generator] = None):
        self.device, self.generator = device, generator

    def randn(self, shape):
        return torch.randn(tuple(shape), device=self.device, generator=self.generator)

    def rand(self, shape):
        return torch.rand(tuple(shape), device=self.device, generator=self.generator)


class InjectedNoise:
    """serves pre-drawn tensors in call order (normal and uniform draws share one sequence, like the patched reference)"""

    def __init__(self, draws, device):
        self.draws, self.device, self.calls = list(draws), device, 0

    def _next(self, shape):
        z = self.draws[self.calls]
        assert tuple(z.shape) == tuple(shape), (self.calls, tuple(z.shape), tuple(shape))
        self.calls += 1
        return z.to(self.device, torch.float32).clone()

    randn = rand = _next


def _normal_log_prob(value, loc, scale):
    """torch.distributions.Normal(loc, scale).log_prob(value)"""
    var = scale ** 2
    return -((value - loc) ** 2) / (2 * var) - torch.log(scale) - math.log(math.sqrt(2 * math.pi))


def sample_loop_energy(diffusion, batch, return_history=False, noise=None):
    """p_sample_loop (ddpm.py:260-340) for the energy-form denoiser: DDPM step on the energy gradient, then ULA / MALA / HMC."""
    den = diffusion.denoise_fn                       # ComposedEBMDenoiseFn
    core = den.model
    dev = core.cuda_device()
    T, P = diffusion.num_timesteps, diffusion.dims[-1][0]
    src = noise if noise is not None else TorchNoise(dev)
    m = batch.mask.bool().to(dev)
    gt = batch.x[:, diffusion.dims[-1][1]:diffusion.dims[-1][2]].to(dev, torch.float32)
    shape = tuple(gt.shape)
    tab = {k: torch.as_tensor(v, device=dev) for k, v in diffusion._tables().items()}
    step_sizes = diffusion.step_sizes.detach().cpu().numpy().astype(np.float32)
    betas = diffusion.betas.detach().to(dev, torch.float32)

    def gradient_function(x, t):                     # ddpm.py:279-283
        g, _ = energy_and_gradient(core, batch, x, t)
        return -g * tab['ula_grad_scale'][t]

    def energy_function(x, t):                       # ddpm.py:285-289 (a scalar for the WHOLE batch, shape [1])
        _, e = energy_and_gradient(core, batch, x, t)
        return (-e * tab['ula_grad_scale'][t]).reshape(1)

    x = 0.5 * src.randn(shape)                       # ddpm.py:273-274
    x[m] = gt[m]
    history = [x.clone()] if return_history else None
    EBM = diffusion.EBM
    sps = diffusion._samples_per_step_table() if EBM and 'ULA' in str(EBM) else None
    for j in reversed(range(T)):
        # p_sample with the energy gradient as the predicted noise (ddpm.py:245-258)
        eps, _ = energy_and_gradient(core, batch, x, j)
        x0 = tab['sqrt_recip_alphas_cumprod'][j] * x - tab['sqrt_recipm1_alphas_cumprod'][j] * eps
        mean = tab['posterior_mean_coef1'][j] * x0 + tab['posterior_mean_coef2'][j] * x
        z = src.randn(shape)
        x = mean + (1 - int(j == 0)) * (0.5 * tab['posterior_log_variance_clipped'][j]).exp() * z
        if EBM and j % max(int(den.ebm_per_steps), 1) == 0:
            if 'ULA' in str(EBM):                                                        # ddpm.py:955-966
                ss = float(step_sizes[j])
                std = float(np.float32(2 * np.float32(ss)) ** np.float32(.5))
                for _ in range(int(sps[j])):
                    x = x + gradient_function(x, j) * ss + src.randn(shape) * std
            elif EBM == 'MALA':                                                          # ddpm.py:1000-1033
                ss = float(step_sizes[j])
                std = float(np.float32(2 * np.float32(ss)) ** np.float32(.5))
                for _ in range(int(diffusion.samples_per_step)):
                    mu = x + gradient_function(x, j) * ss
                    scale = torch.ones_like(x) * std
                    x_hat = mu + src.randn(shape) * std
                    logp_x, logp_x_hat = energy_function(x, j), energy_function(x_hat, j)
                    logp_reverse = _normal_log_prob(x, mu, scale).sum(1)
                    logp_forward = _normal_log_prob(x_hat, mu, scale).sum(1)
                    logp_accept = logp_x_hat - logp_x + logp_reverse - logp_forward
                    accept = (src.rand((shape[0],)) < torch.exp(logp_accept)).float()
                    x = accept[:, None] * x_hat + (1 - accept[:, None]) * x
            elif EBM == 'HMC':                                                           # ddpm.py:293-318, 1036-1126
                n_samples, n_leapfrog, damping = 4, 2, 0
                mass = 9 * betas
                m_t = mass[j]
                v = src.randn(shape) * m_t
                for i in range(n_samples):
                    eps_i = src.randn(shape)
                    v_prime = v * damping + float(np.sqrt(1. - damping ** 2)) * eps_i * m_t
                    # leapfrog (ddpm.py:917-937): step size, mass AND the timestep handed to the gradient are indexed by i (sic, :1071-1079)
                    h, m_i = float(step_sizes[i]), mass[i]
                    x_k, v_k = x, v_prime
                    for _ in range(n_leapfrog):
                        v_k = v_k + 0.5 * h * gradient_function(x_k, i)
                        x_k = x_k + h * v_k / (m_i ** 2.)
                        v_k = v_k + 0.5 * h * gradient_function(x_k, i)
                    zero, sc = torch.zeros_like(x), torch.ones_like(x) * m_t
                    logp_v_p = _normal_log_prob(v_prime, zero, sc).sum(1)
                    logp_v = _normal_log_prob(v_k, zero, sc).sum(1)
                    logp_accept = (energy_function(x_k, j) + logp_v) - (energy_function(x, j) + logp_v_p)
                    accept = (src.rand((shape[0],)) < torch.exp(logp_accept)).float()
                    x = accept[:, None] * x_k + (1 - accept[:, None]) * x
                    v = accept[:, None] * v_k + (1 - accept[:, None]) * v_prime
            else:
                raise NotImplementedError(f'EBM={EBM!r}')
        x[m] = gt[m]                                                                      # ddpm.py:334
        if return_history:
            history.append(x.clone())
    return (x, history) if return_history else x
