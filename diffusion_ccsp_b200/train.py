"""Training step on the GPU: host-side glue between the nn.Parameters of ConstraintDiffuser and
`ccsp_train_step` / `ccsp_adam_step` (C ABI, hand-written FP32 kernels).

  * TrainGraph        compiled training batch (replaces the per-call graph handling of denoise_fn.py:466-521)
  * diffusion_loss    loss of GaussianDiffusion.p_losses (networks/ddpm.py:363-385) as a torch scalar whose `.backward()`
                      delivers the gradients the kernels computed (so `loss.backward()` / gradient accumulation of
                      Trainer.train, ddpm.py:533-534, work unchanged)
  * Adam              torch.optim.Adam-shaped optimiser over `ccsp_adam_step` (ddpm.py:466, 542-543)

There is no CPU fallback: parameters must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _abi

ENCODERS = (('geom_encoder', 'geom'), ('grasp_encoder', 'grasp'), ('pose_encoder', 'pose'))


class TrainDims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('hidden_dim', 'geom_dim', 'pose_dim', 'grasp_dim', 'num_types', 'normalize',
                                         'row_width', 'pose_begin', 'grasp_begin')]


class Params(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        'geom_w0', 'geom_b0', 'geom_w2', 'geom_b2', 'grasp_w0', 'grasp_b0', 'grasp_w2', 'grasp_b2',
        'pose_w0', 'pose_b0', 'pose_w2', 'pose_b2', 'dec_w0', 'dec_b0', 'dec_w2', 'dec_b2',
        'time_w1', 'time_b1', 'time_w3', 'time_b3')] + [('mlp_w', C.POINTER(C.c_void_p)), ('mlp_b', C.POINTER(C.c_void_p))]


def _declare(lib):
    if getattr(lib, '_ccsp_train_declared', False):
        return lib
    lib.ccsp_train_graph_create.argtypes = [C.POINTER(TrainDims), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_int64, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.ccsp_train_graph_destroy.argtypes = [C.c_void_p]
    lib.ccsp_train_graph_destroy.restype = None
    lib.ccsp_train_graph_num_edges_of_type.argtypes = [C.c_void_p, C.c_int32]
    lib.ccsp_train_graph_num_edges_of_type.restype = C.c_int64
    lib.ccsp_train_step.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Params), C.c_int32, C.c_float, C.c_float, C.c_void_p,
                                    C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ccsp_adam_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_float,
                                   C.c_float, C.c_float, C.c_void_p]
    lib._ccsp_train_declared = True
    return lib


def param_names(denoise_fn) -> List[str]:
    """state_dict-style names of every trainable tensor, in the order the autograd Function receives them"""
    names = []
    for enc, _ in ENCODERS:
        if hasattr(denoise_fn, enc):
            names += [f'{enc}.0.weight', f'{enc}.0.bias', f'{enc}.2.weight', f'{enc}.2.bias']
    names += ['pose_decoder.0.weight', 'pose_decoder.0.bias', 'pose_decoder.2.weight', 'pose_decoder.2.bias',
              'time_mlp.1.weight', 'time_mlp.1.bias', 'time_mlp.3.weight', 'time_mlp.3.bias']
    for c in range(len(denoise_fn.mlps)):
        names += [f'mlps.{c}.0.weight', f'mlps.{c}.0.bias']
    return names


def _pack(tensors: dict, num_types: int):
    """name -> CUDA tensor  =>  (CcspParams, keep-alive list)"""
    s = Params()
    keep = []

    def ptr(name):
        t = tensors[name]
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), name
        return t.data_ptr()

    for enc, pre in ENCODERS:
        if f'{enc}.0.weight' in tensors:
            setattr(s, pre + '_w0', ptr(f'{enc}.0.weight')); setattr(s, pre + '_b0', ptr(f'{enc}.0.bias'))
            setattr(s, pre + '_w2', ptr(f'{enc}.2.weight')); setattr(s, pre + '_b2', ptr(f'{enc}.2.bias'))
    s.dec_w0, s.dec_b0 = ptr('pose_decoder.0.weight'), ptr('pose_decoder.0.bias')
    s.dec_w2, s.dec_b2 = ptr('pose_decoder.2.weight'), ptr('pose_decoder.2.bias')
    s.time_w1, s.time_b1 = ptr('time_mlp.1.weight'), ptr('time_mlp.1.bias')
    s.time_w3, s.time_b3 = ptr('time_mlp.3.weight'), ptr('time_mlp.3.bias')
    ws = (C.c_void_p * num_types)(*[ptr(f'mlps.{c}.0.weight') for c in range(num_types)])
    bs = (C.c_void_p * num_types)(*[ptr(f'mlps.{c}.0.bias') for c in range(num_types)])
    s.mlp_w, s.mlp_b = ws, bs
    keep += [ws, bs]
    return s, keep


class TrainGraph:
    """Owns a CcspTrainGraph handle: one training batch compiled for the loss + gradient kernels."""

    def __init__(self, denoise_fn, batch, device):
        self._lib = _declare(_abi.load_library())
        if not torch.cuda.is_available():
            raise _abi.CcspError('CUDA device required: the training step has no CPU fallback')
        self.device = torch.device(device)
        dims = denoise_fn.dims
        robot = 'robot' in denoise_fn.input_mode
        d = TrainDims()
        d.hidden_dim = 256
        d.geom_dim, d.pose_dim = dims[0][0], dims[-1][0]
        d.grasp_dim = dims[1][0] if robot else 0
        d.num_types, d.normalize = len(denoise_fn.mlps), int(bool(denoise_fn.normalize))
        d.row_width, d.pose_begin = int(batch.x.shape[1]), dims[-1][1]
        d.grasp_begin = dims[1][1] if robot else 0
        x = batch.x.detach().to('cpu', torch.float32).contiguous()
        ei = batch.edge_index.detach().to('cpu', torch.int64).contiguous()
        ea = batch.edge_attr.detach().to('cpu', torch.float32).contiguous()
        mk = batch.mask.detach().to('cpu', torch.int8).contiguous()
        self.n, self.P, self.num_types = int(x.shape[0]), int(d.pose_dim), int(d.num_types)
        E = int(ei.shape[1]) if ei.numel() else 0
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _abi.check(self._lib.ccsp_train_graph_create(C.byref(d), x.data_ptr(), self.n, ei.data_ptr() if E else None,
                                                         ea.data_ptr() if E else None, mk.data_ptr(), E,
                                                         _abi.current_stream_ptr(self.device), C.byref(h)),
                       'ccsp_train_graph_create')
        self._h = h
        self.mask = mk.bool().to(self.device)
        self.edges_of_type = [int(self._lib.ccsp_train_graph_num_edges_of_type(h, c)) for c in range(self.num_types)]

    def step(self, weights: dict, grads: dict, t: int, sqrt_ac: float, sqrt_1mac: float, noise: torch.Tensor, loss_l1: bool,
             grad_scale: float, loss_out: torch.Tensor, recon: Optional[torch.Tensor] = None):
        w, k1 = _pack(weights, self.num_types)
        g, k2 = _pack(grads, self.num_types)
        assert noise.is_cuda and noise.dtype == torch.float32 and noise.is_contiguous() and tuple(noise.shape) == (self.n, self.P)
        with torch.cuda.device(self.device):
            _abi.check(self._lib.ccsp_train_step(self._h, C.byref(w), C.byref(g), int(t), float(sqrt_ac), float(sqrt_1mac),
                                                 noise.data_ptr(), int(bool(loss_l1)), float(grad_scale), loss_out.data_ptr(),
                                                 recon.data_ptr() if recon is not None else None,
                                                 _abi.current_stream_ptr(self.device)), 'ccsp_train_step')
        del k1, k2

    def close(self):
        if getattr(self, '_h', None):
            with torch.cuda.device(self.device):
                self._lib.ccsp_train_graph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _DiffusionLoss(torch.autograd.Function):
    """loss = p_losses(...) computed by ccsp_train_step together with its gradients; backward hands them to autograd."""

    @staticmethod
    def forward(ctx, graph: TrainGraph, names, t, sqrt_ac, sqrt_1mac, noise, loss_l1, recon, *params):
        weights = {n: p.detach() for n, p in zip(names, params)}
        grads = {n: torch.empty_like(p) for n, p in weights.items()}
        loss = torch.empty((), dtype=torch.float32, device=graph.device)
        graph.step(weights, grads, t, sqrt_ac, sqrt_1mac, noise, loss_l1, 1.0, loss, recon)
        out = []
        for n in names:
            # a constraint type without edges never enters the reference's autograd graph: its grads stay None (denoise_fn.py:514-515)
            if n.startswith('mlps.') and graph.edges_of_type[int(n.split('.')[1])] == 0:
                out.append(None)
            else:
                out.append(grads[n])
        ctx.grads = out
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        live = [g for g in ctx.grads if g is not None]
        torch._foreach_mul_(live, grad_out)           # one multi-tensor launch; the buffers belong to this call
        return (None,) * 8 + tuple(ctx.grads)


def diffusion_loss(denoise_fn, graph: TrainGraph, t: int, sqrt_ac: float, sqrt_1mac: float, noise: torch.Tensor,
                   loss_type: str = 'l2', recon: Optional[torch.Tensor] = None) -> torch.Tensor:
    if loss_type not in ('l1', 'l2'):
        raise NotImplementedError(loss_type)                                         # ddpm.py:384-385
    names = param_names(denoise_fn)
    sd = dict(denoise_fn.named_parameters())
    params = [sd[n] for n in names]
    for n, p in zip(names, params):
        if not p.is_cuda:
            raise _abi.CcspError(f'parameter {n} is on {p.device}: move the model to the CUDA device first (.to(device) / .cuda()); '
                                 'the training step has no CPU fallback')
    return _DiffusionLoss.apply(graph, names, int(t), float(sqrt_ac), float(sqrt_1mac), noise, loss_type == 'l1', recon, *params)


class Adam:
    """torch.optim.Adam-shaped optimiser (defaults of ddpm.py:466: betas (0.9, 0.999), eps 1e-8, no weight decay) over
    ccsp_adam_step.  Parameters whose `.grad` is None are skipped entirely, like torch does."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, on_step=None):
        self._lib = _declare(_abi.load_library())
        self._on_step = on_step      # e.g. ConstraintDiffuser.mark_weights_dirty: the update goes through raw device pointers
        self.params = [p for p in params]
        self.param_groups = [dict(lr=lr, betas=betas, eps=eps, params=self.params)]
        self.state = {}

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self):
        g = self.param_groups[0]
        for p in self.params:
            if p.grad is None:
                continue
            if not p.is_cuda:
                raise _abi.CcspError('Adam: parameters must live on a CUDA device (no CPU fallback)')
            st = self.state.setdefault(p, dict(step=0, exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p)))
            st['step'] += 1
            grad = p.grad.contiguous()
            with torch.cuda.device(p.device):
                _abi.check(self._lib.ccsp_adam_step(p.data_ptr(), grad.data_ptr(), st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr(),
                                                    p.numel(), st['step'], float(g['lr']), float(g['betas'][0]), float(g['betas'][1]),
                                                    float(g['eps']), _abi.current_stream_ptr(p.device)), 'ccsp_adam_step')
        if self._on_step is not None:
            self._on_step()

    def state_dict(self):
        return dict(state={i: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in self.state.get(p, {}).items()}
                           for i, p in enumerate(self.params)},
                    param_groups=[{k: v for k, v in self.param_groups[0].items() if k != 'params'}])
