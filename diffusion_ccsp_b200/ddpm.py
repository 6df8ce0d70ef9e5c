"""GaussianDiffusion — host-side mirror of networks/ddpm.py:168-390 (sampling half) over the CUDA path.

Same constructor arguments, buffers (so `load_state_dict` of a reference checkpoint works), attributes
and `sample` / `p_sample_loop` / `p_sample` / `p_mean_variance` signatures.  The T x (1+K) loop of
p_sample_loop (ddpm.py:325-336) and AnnealedULASampler.sample_step (ddpm.py:955-966) runs inside
`ccsp_sample` (C ABI): no Python per-timestep or per-edge work.
"""
from __future__ import annotations

import time
from typing import Optional

import numpy as np
import torch
import torch.nn as nn


def cosine_beta_schedule(timesteps, s=0.008):
    """networks/ddpm.py:152-162."""
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    alphas_cumprod = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    alphas_cumprod = alphas_cumprod / alphas_cumprod[0]
    betas = 1 - (alphas_cumprod[1:] / alphas_cumprod[:-1])
    return np.clip(betas, a_min=0, a_max=0.999)


def _resolve_step_sizes(expr, betas: torch.Tensor) -> torch.Tensor:
    """The reference `eval`s a Python expression in the constructor (ddpm.py:207, default
    '2*self.betas').  We accept the same strings but evaluate them in a closed namespace once, or a
    ready tensor/array."""
    if isinstance(expr, str):
        class _Self:
            pass
        s = _Self()
        s.betas = betas
        return torch.as_tensor(eval(expr, {'__builtins__': {}}, {'self': s, 'torch': torch, 'np': np}),
                               dtype=torch.float32)
    return torch.as_tensor(expr, dtype=torch.float32)


class GaussianDiffusion(nn.Module):
    def __init__(self, denoise_fn, timesteps=100, loss_type='l2', EBM=False, betas=None,
                 samples_per_step=10, step_sizes='2*self.betas'):
        super().__init__()
        if betas is not None:
            betas = betas.detach().cpu().numpy() if isinstance(betas, torch.Tensor) else np.asarray(betas)
        else:
            betas = cosine_beta_schedule(timesteps)
        self._betas = betas
        alphas = 1. - betas
        alphas_cumprod = np.cumprod(alphas, axis=0)
        alphas_cumprod_prev = np.append(1., alphas_cumprod[:-1])
        timesteps, = betas.shape
        self.denoise_fn = denoise_fn
        self.device = denoise_fn.device
        self.dims = denoise_fn.dims
        self.input_mode = denoise_fn.input_mode
        self.num_timesteps = int(timesteps)
        self.loss_type = loss_type
        if EBM not in (False, None, 'ULA', 'ULA+', 'MALA', 'HMC'):
            raise NotImplementedError(f'EBM={EBM!r}: choose False / ULA / ULA+ / MALA / HMC (train_utils.py:92)')
        if EBM in ('MALA', 'HMC') and not getattr(denoise_fn, 'energy_wrapper', False):
            raise ValueError(f'EBM={EBM!r} needs the energy form: wrap the denoiser in ComposedEBMDenoiseFn with '
                             'energy_wrapper=True (train_utils.py:115-116, 283-284)')
        self.EBM = EBM

        def to_torch(a):
            return torch.tensor(a, dtype=torch.float32)

        self.register_buffer('betas', to_torch(betas))
        self.register_buffer('alphas_cumprod', to_torch(alphas_cumprod))
        self.register_buffer('alphas_cumprod_prev', to_torch(alphas_cumprod_prev))
        self.samples_per_step = samples_per_step
        self.step_sizes = _resolve_step_sizes(step_sizes, self.betas)                       # ddpm.py:207
        self.register_buffer('sqrt_alphas_cumprod', to_torch(np.sqrt(alphas_cumprod)))
        self.register_buffer('sqrt_one_minus_alphas_cumprod', to_torch(np.sqrt(1. - alphas_cumprod)))
        self.register_buffer('log_one_minus_alphas_cumprod', to_torch(np.log(1. - alphas_cumprod)))
        self.register_buffer('sqrt_recip_alphas_cumprod', to_torch(np.sqrt(1. / alphas_cumprod)))
        self.register_buffer('sqrt_recipm1_alphas_cumprod', to_torch(np.sqrt(1. / alphas_cumprod - 1)))
        self._sqrt_recipm1_alphas_cumprod_custom = to_torch(np.sqrt(1. / (1 - alphas_cumprod)))   # ddpm.py:215
        posterior_variance = betas * (1. - alphas_cumprod_prev) / (1. - alphas_cumprod)
        self.register_buffer('posterior_variance', to_torch(posterior_variance))
        self.register_buffer('posterior_log_variance_clipped', to_torch(np.log(np.maximum(posterior_variance, 1e-20))))
        self.register_buffer('posterior_mean_coef1', to_torch(betas * np.sqrt(alphas_cumprod_prev) / (1. - alphas_cumprod)))
        self.register_buffer('posterior_mean_coef2', to_torch((1. - alphas_cumprod_prev) * np.sqrt(alphas) / (1. - alphas_cumprod)))
        self.sample_loop_time = []

    # ------------------------------------------------------------------------------------------
    def _tables(self):
        g = lambda t: t.detach().cpu().numpy().astype(np.float32)
        return {
            'sqrt_recip_alphas_cumprod': g(self.sqrt_recip_alphas_cumprod),
            'sqrt_recipm1_alphas_cumprod': g(self.sqrt_recipm1_alphas_cumprod),
            'posterior_mean_coef1': g(self.posterior_mean_coef1),
            'posterior_mean_coef2': g(self.posterior_mean_coef2),
            'posterior_log_variance_clipped': g(self.posterior_log_variance_clipped),
            'ula_grad_scale': g(self._sqrt_recipm1_alphas_cumprod_custom),
            'step_sizes': g(self.step_sizes),
        }

    def _samples_per_step_table(self) -> Optional[np.ndarray]:
        """ULA steps per timestep (ddpm.py:294-302); 'ULA+' = 4/8/12/16 per quarter, low t -> high t."""
        T = self.num_timesteps
        if not self.EBM:
            return None
        if self.EBM == 'ULA+':
            n = T // 4
            tab = np.array([4] * n + [8] * n + [12] * n + [16] * n, dtype=np.int32)
            if tab.shape[0] != T:
                raise ValueError('ULA+ needs timesteps divisible by 4 (ddpm.py:298-299)')
            return tab
        if isinstance(self.samples_per_step, int):
            return np.full((T,), self.samples_per_step, dtype=np.int32)
        return np.asarray(self.samples_per_step, dtype=np.int32).reshape(T)

    def num_noise_draws(self) -> int:
        sps = self._samples_per_step_table()
        per = max(int(self.denoise_fn.ebm_per_steps), 1)
        k = 0 if sps is None else int(sum(int(sps[j]) for j in range(self.num_timesteps) if j % per == 0))
        return 1 + self.num_timesteps + k

    # ------------------------------------------------------------------------------------------
    def p_mean_variance(self, batch, features, t, clip_denoised: bool = False, **kwargs):
        """ddpm.py:245-251 (single-step API kept for callers that drive the loop themselves)."""
        kwargs.pop('noise', None)
        eps = self.denoise_fn(features, batch, t, eval=True, **kwargs)
        dev = eps.device
        ti = int(t.reshape(-1)[0].item()) if torch.is_tensor(t) else int(t)
        f = features.to(dev)
        x0 = self.sqrt_recip_alphas_cumprod[ti].to(dev) * f - self.sqrt_recipm1_alphas_cumprod[ti].to(dev) * eps
        if clip_denoised:
            x0.clamp_(-1., 1.)
        mean = self.posterior_mean_coef1[ti].to(dev) * x0 + self.posterior_mean_coef2[ti].to(dev) * f
        return mean, self.posterior_variance[ti].to(dev), self.posterior_log_variance_clipped[ti].to(dev)

    def p_sample(self, batch, all_features, t, repeat_noise=False, noise=None, **kwargs):
        """ddpm.py:253-258."""
        mean, _, logvar = self.p_mean_variance(batch, all_features, t, **kwargs)
        z = torch.randn(all_features.shape, device=mean.device) if noise is None else noise.to(mean.device)
        ti = int(t.reshape(-1)[0].item()) if torch.is_tensor(t) else int(t)
        return mean + (1 - int(ti == 0)) * (0.5 * logvar).exp() * z

    @torch.no_grad()
    def p_sample_loop(self, batch, return_history=False, noise=None, x_init=None, seed=None,
                      node_offset=0, **kwargs):
        """ddpm.py:260-340.  Extra keyword arguments (not in the reference):
             noise  [1+T(1+K), n, P]  injected Gaussian draws in the reference's draw order (parity runs);
             x_init [n, P]            start state instead of 0.5*randn;
             seed / node_offset       Philox stream when `noise` is None (node_offset = first global node
                                      id of this shard, so sharded runs reproduce the single-GPU stream).
        Returns poses [n,P] on the CUDA device (and a list of T+1 tensors when return_history)."""
        assert not self.training                                                    # ddpm.py:328
        den = self.denoise_fn
        if getattr(den, 'energy_wrapper', False):       # energy form: host-driven MCMC around ccsp_energy_grad (ebm.py)
            from .ebm import InjectedNoise, sample_loop_energy
            if noise is not None and not hasattr(noise, 'randn'):
                noise = InjectedNoise(noise, den.model.cuda_device())
            return sample_loop_energy(self, batch, return_history=return_history, noise=noise)
        plan = den.plan_for(batch, verify_content=True)
        dev = plan.model.device
        n, P, T = plan.n, self.dims[-1][0], self.num_timesteps
        out = torch.empty((n, P), dtype=torch.float32, device=dev)
        hist = torch.empty((T + 1, n, P), dtype=torch.float32, device=dev) if return_history else None
        if noise is not None:
            noise = noise.to(dev, torch.float32).contiguous()
            if tuple(noise.shape) != (self.num_noise_draws(), n, P):
                raise ValueError(f'noise must be [{self.num_noise_draws()}, {n}, {P}], got {tuple(noise.shape)}')
        if x_init is not None:
            x_init = x_init.to(dev, torch.float32).contiguous()
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if noise is None else 0
        plan.sample(self._tables(), self._samples_per_step_table(), den.ebm_per_steps, out, hist,
                    noise=noise, x_init=x_init, seed=seed, node_offset=node_offset)
        if return_history:
            return out, list(hist.unbind(0))
        return out

    @torch.no_grad()
    def sample(self, batch, **kwargs):
        """ddpm.py:342-351 (wall-clock bookkeeping included; synchronises to make the time meaningful)."""
        kwargs.pop('debug', None)
        start = time.time()
        outputs = self.p_sample_loop(batch, **kwargs)
        torch.cuda.synchronize(outputs[0].device if isinstance(outputs, tuple) else outputs.device)
        passed = time.time() - start
        self.sample_loop_time.append(passed)
        if len(self.sample_loop_time) > 10:
            self.sample_loop_time.pop(0)
        return outputs

    # ------------------------------------------------------------------------------------------
    # training half (ddpm.py:353-389): loss and gradients come from ccsp_train_step (CUDA)
    # ------------------------------------------------------------------------------------------
    def q_sample(self, x_start, mask, t, noise=None):
        """ddpm.py:353-361 (kept for callers; p_losses fuses it into the training kernels)."""
        if noise is None:
            noise = conditional_noise(x_start, mask)
        ti = int(t.reshape(-1)[0].item()) if torch.is_tensor(t) else int(t)
        dev = x_start.device
        sample = self.sqrt_alphas_cumprod[ti].to(dev) * x_start + self.sqrt_one_minus_alphas_cumprod[ti].to(dev) * noise
        sample[mask.bool()] = x_start[mask.bool()]
        return sample

    def p_losses(self, batch, t, noise=None, **kwargs):
        """ddpm.py:363-385: q_sample with masked noise, denoiser, l1 / l2 loss — one C call (loss + every gradient); the
        returned scalar's .backward() delivers the gradients to the nn.Parameters.  `recon=` (a [n,P] CUDA tensor, extra
        keyword) receives the denoiser output."""
        from .train import diffusion_loss
        den = self.denoise_fn
        graph = den.train_graph_for(batch)
        dev = graph.device
        ti = int(t.reshape(-1)[0].item()) if torch.is_tensor(t) else int(t)
        if not 0 <= ti < self.num_timesteps:
            raise IndexError(f't={ti} outside the schedule [0, {self.num_timesteps})')
        if noise is None:
            noise = torch.randn((graph.n, graph.P), dtype=torch.float32, device=dev)    # conditional_noise (ddpm.py:114-117)
            noise[graph.mask] = 0
        noise = noise.detach().to(dev, torch.float32).contiguous()
        recon = kwargs.pop('recon', None)
        loss = diffusion_loss(den, graph, ti, float(self.sqrt_alphas_cumprod[ti]), float(self.sqrt_one_minus_alphas_cumprod[ti]),
                              noise, self.loss_type, recon)
        if kwargs['debug']:                                                             # read unguarded in the reference (ddpm.py:374)
            print(f'[p_losses] t={ti} loss={float(loss):.6f}')
        return loss

    def forward(self, batch, **kwargs):
        """ddpm.py:387-389: ONE random timestep per batch."""
        t = torch.randint(0, self.num_timesteps, (1,)).long()
        return self.p_losses(batch, t, **kwargs)


def conditional_noise(x, mask):
    """ddpm.py:114-117."""
    noise = torch.randn_like(x)
    noise[mask.bool()] = 0
    return noise
