"""ctypes binding of libccsp_b200.so (include/ccsp_b200.h).

There is deliberately no fallback: if the library is missing or a call fails this module raises —
the CUDA path is the product, the oracle under oracle/ is only a checker.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np
import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, 'lib', 'libccsp_b200.so')

MATH_FP32, MATH_TF32X3, MATH_BF16X3, MATH_TF32, MATH_BF16 = 0, 1, 2, 3, 4
MATH_NAMES = {'fp32': MATH_FP32, 'tf32x3': MATH_TF32X3, 'bf16x3': MATH_BF16X3, 'tf32': MATH_TF32, 'bf16': MATH_BF16}

EXPORTS = [
    'ccsp_last_error', 'ccsp_debug_trap_info', 'ccsp_debug_persist_trace', 'ccsp_debug_chain_cuts', 'ccsp_abi_version', 'ccsp_launch_count', 'ccsp_reset_launch_count',
    'ccsp_model_create', 'ccsp_model_destroy', 'ccsp_model_set_math', 'ccsp_model_get_math',
    'ccsp_plan_create', 'ccsp_plan_destroy', 'ccsp_plan_num_nodes', 'ccsp_plan_num_edges',
    'ccsp_plan_num_edge_rows', 'ccsp_denoise', 'ccsp_sample',
    'ccsp_plan_set_timing', 'ccsp_plan_get_timing', 'ccsp_plan_h2d_bytes',
    'ccsp_check_solved',
    'ccsp_train_graph_create', 'ccsp_train_graph_destroy', 'ccsp_train_graph_num_edges_of_type', 'ccsp_train_step',
    'ccsp_adam_step', 'ccsp_energy_grad',
]


class CcspError(RuntimeError):
    pass


class ModelDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('hidden_dim', 'geom_dim', 'pose_dim', 'grasp_dim', 'num_types', 'normalize')] + \
               [(n, C.c_void_p) for n in (
                   'geom_w0', 'geom_b0', 'geom_w2', 'geom_b2',
                   'grasp_w0', 'grasp_b0', 'grasp_w2', 'grasp_b2',
                   'pose_w0', 'pose_b0', 'pose_w2', 'pose_b2',
                   'dec_w0', 'dec_b0', 'dec_w2', 'dec_b2',
                   'time_w1', 'time_b1', 'time_w3', 'time_b3')] + \
               [('mlp_w', C.POINTER(C.c_void_p)), ('mlp_b', C.POINTER(C.c_void_p))]


class Schedule(C.Structure):
    _fields_ = [('T', C.c_int32)] + \
               [(n, C.c_void_p) for n in (
                   'sqrt_recip_alphas_cumprod', 'sqrt_recipm1_alphas_cumprod', 'posterior_mean_coef1',
                   'posterior_mean_coef2', 'posterior_log_variance_clipped', 'ula_grad_scale', 'step_sizes',
                   'samples_per_step')] + \
               [('ebm_per_steps', C.c_int32)]


class Timing(C.Structure):
    _fields_ = [('samples', C.c_int64), ('ms_edge_l1', C.c_double), ('ms_edge_dec', C.c_double), ('ms_node', C.c_double)]


class Noise(C.Structure):
    _fields_ = [('x_init', C.c_void_p), ('noise', C.c_void_p), ('seed', C.c_uint64), ('node_offset', C.c_uint64)]


_lib = None


def load_library(path: Optional[str] = None):
    """dlopen the C-ABI library and declare the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise CcspError(
            f'{path} not found: build it with `python -m diffusion_ccsp_b200.build` '
            '(or __graft_entry__.build()).  There is no CPU fallback.')
    lib = C.CDLL(path)
    lib.ccsp_last_error.restype = C.c_char_p
    lib.ccsp_abi_version.restype = C.c_int
    lib.ccsp_launch_count.restype = C.c_uint64
    lib.ccsp_reset_launch_count.restype = None
    lib.ccsp_model_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(C.c_void_p)]
    lib.ccsp_model_destroy.argtypes = [C.c_void_p]
    lib.ccsp_model_destroy.restype = None
    lib.ccsp_model_set_math.argtypes = [C.c_void_p, C.c_int]
    lib.ccsp_model_get_math.argtypes = [C.c_void_p]
    lib.ccsp_plan_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.ccsp_plan_destroy.argtypes = [C.c_void_p]
    lib.ccsp_plan_destroy.restype = None
    lib.ccsp_plan_set_timing.argtypes = [C.c_void_p, C.c_int32]
    lib.ccsp_plan_get_timing.argtypes = [C.c_void_p, C.POINTER(Timing)]
    for f in ('ccsp_plan_num_nodes', 'ccsp_plan_num_edges', 'ccsp_plan_num_edge_rows', 'ccsp_plan_h2d_bytes'):
        getattr(lib, f).argtypes = [C.c_void_p]
        getattr(lib, f).restype = C.c_int64
    lib.ccsp_denoise.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.ccsp_sample.argtypes = [C.c_void_p, C.POINTER(Schedule), C.POINTER(Noise), C.c_void_p, C.c_void_p, C.c_void_p]
    if path == LIB_PATH:
        _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load_library().ccsp_last_error().decode(errors='replace')
        raise CcspError(f'{what} failed (status {rc}): {msg}')


def _f32_host(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to('cpu', torch.float32).contiguous()


def current_stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Model:
    """Owns a CcspModel handle packed from reference-layout weights (a state_dict-like mapping with the
    `denoise_fn.`-less keys of SURVEY.md §8b)."""

    def __init__(self, weights, dims, num_types: int, normalize: bool, device: torch.device, math: str = 'fp32'):
        lib = load_library()
        if not torch.cuda.is_available():
            raise CcspError('CUDA device required: diffusion_ccsp_b200 has no CPU fallback')
        self.device = torch.device(device)
        robot = len(dims) == 3
        keep = []

        def ptr(name):
            t = _f32_host(weights[name])
            keep.append(t)
            return t.data_ptr()

        d = ModelDesc()
        d.hidden_dim = 256
        d.geom_dim, d.pose_dim = dims[0][0], dims[-1][0]
        d.grasp_dim = dims[1][0] if robot else 0
        d.num_types, d.normalize = num_types, int(bool(normalize))
        for enc, pre in (('geom_encoder', 'geom'), ('pose_encoder', 'pose')) + ((('grasp_encoder', 'grasp'),) if robot else ()):
            setattr(d, pre + '_w0', ptr(f'{enc}.0.weight')); setattr(d, pre + '_b0', ptr(f'{enc}.0.bias'))
            setattr(d, pre + '_w2', ptr(f'{enc}.2.weight')); setattr(d, pre + '_b2', ptr(f'{enc}.2.bias'))
        d.dec_w0, d.dec_b0 = ptr('pose_decoder.0.weight'), ptr('pose_decoder.0.bias')
        d.dec_w2, d.dec_b2 = ptr('pose_decoder.2.weight'), ptr('pose_decoder.2.bias')
        d.time_w1, d.time_b1 = ptr('time_mlp.1.weight'), ptr('time_mlp.1.bias')
        d.time_w3, d.time_b3 = ptr('time_mlp.3.weight'), ptr('time_mlp.3.bias')
        ws = (C.c_void_p * num_types)(*[ptr(f'mlps.{c}.0.weight') for c in range(num_types)])
        bs = (C.c_void_p * num_types)(*[ptr(f'mlps.{c}.0.bias') for c in range(num_types)])
        d.mlp_w, d.mlp_b = ws, bs
        k_in = 256 * (6 if robot else 5)
        assert tuple(weights['mlps.0.0.weight'].shape) == (512, k_in), weights['mlps.0.0.weight'].shape
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.ccsp_model_create(C.byref(d), C.byref(h)), 'ccsp_model_create')
        self._h = h
        self._lib = lib
        self.set_math(math)

    def set_math(self, math: str):
        check(self._lib.ccsp_model_set_math(self._h, MATH_NAMES[math]), f'ccsp_model_set_math({math})')
        self.math = math

    def close(self):
        if getattr(self, '_h', None):
            self._lib.ccsp_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Plan:
    """Owns a CcspPlan handle: the HBM-resident compiled form of one scene batch."""

    def __init__(self, model: Model, x: torch.Tensor, edge_index: torch.Tensor, edge_attr: torch.Tensor,
                 mask: torch.Tensor, pose_begin: int, grasp_begin: int = 0):
        self.model = model
        self._lib = model._lib
        x = _f32_host(x)
        ei = edge_index.detach().to('cpu', torch.int64).contiguous()
        ea = _f32_host(edge_attr)
        mk = mask.detach().to('cpu', torch.int8).contiguous()
        n, F = x.shape
        E = ei.shape[1] if ei.numel() else 0
        h = C.c_void_p()
        with torch.cuda.device(model.device):
            check(self._lib.ccsp_plan_create(model._h, x.data_ptr(), n, F, ei.data_ptr() if E else None,
                                             ea.data_ptr() if E else None, mk.data_ptr(), E, pose_begin, grasp_begin,
                                             current_stream_ptr(model.device), C.byref(h)), 'ccsp_plan_create')
        self._h = h
        self.n, self.E, self.P = n, E, None
        self.edge_rows = int(self._lib.ccsp_plan_num_edge_rows(h))
        self.h2d_bytes = int(self._lib.ccsp_plan_h2d_bytes(h))

    def set_timing(self, stride: int):
        check(self._lib.ccsp_plan_set_timing(self._h, int(stride)), 'ccsp_plan_set_timing')

    def get_timing(self) -> dict:
        t = Timing()
        check(self._lib.ccsp_plan_get_timing(self._h, C.byref(t)), 'ccsp_plan_get_timing')
        return dict(samples=int(t.samples), ms_edge_l1=t.ms_edge_l1, ms_edge_dec=t.ms_edge_dec, ms_node=t.ms_node)

    def denoise(self, poses: torch.Tensor, t: int, out: torch.Tensor):
        dev = self.model.device
        assert poses.is_cuda and out.is_cuda and poses.dtype == torch.float32 and poses.is_contiguous()
        with torch.cuda.device(dev):
            check(self._lib.ccsp_denoise(self._h, poses.data_ptr(), int(t), out.data_ptr(), current_stream_ptr(dev)),
                  'ccsp_denoise')
        return out

    def sample(self, tables: dict, samples_per_step: Optional[np.ndarray], ebm_per_steps: int,
               out: torch.Tensor, history: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
               x_init: Optional[torch.Tensor] = None, seed: int = 0, node_offset: int = 0):
        """tables: name -> contiguous float32 numpy arrays [T] (host)."""
        dev = self.model.device
        s = Schedule()
        T = int(tables['sqrt_recip_alphas_cumprod'].shape[0])
        s.T = T
        keep = []

        def hp(a, dtype=np.float32):
            a = np.ascontiguousarray(a, dtype=dtype)
            assert a.shape == (T,), a.shape
            keep.append(a)
            return a.ctypes.data

        for k in ('sqrt_recip_alphas_cumprod', 'sqrt_recipm1_alphas_cumprod', 'posterior_mean_coef1',
                  'posterior_mean_coef2', 'posterior_log_variance_clipped'):
            setattr(s, k, hp(tables[k]))
        if samples_per_step is not None:
            s.ula_grad_scale = hp(tables['ula_grad_scale'])
            s.step_sizes = hp(tables['step_sizes'])
            s.samples_per_step = hp(samples_per_step, np.int32)
        s.ebm_per_steps = int(ebm_per_steps)
        nz = Noise()
        nz.x_init = x_init.data_ptr() if x_init is not None else None
        nz.noise = noise.data_ptr() if noise is not None else None
        nz.seed, nz.node_offset = int(seed) & (2 ** 64 - 1), int(node_offset)
        for t_ in (out, history, noise, x_init):
            if t_ is not None:
                assert t_.is_cuda and t_.dtype == torch.float32 and t_.is_contiguous()
        with torch.cuda.device(dev):
            check(self._lib.ccsp_sample(self._h, C.byref(s), C.byref(nz), out.data_ptr(),
                                        history.data_ptr() if history is not None else None,
                                        current_stream_ptr(dev)), 'ccsp_sample')
        return out

    def close(self):
        if getattr(self, '_h', None):
            with torch.cuda.device(self.model.device):      # destroy synchronises the plan's device, not the caller's
                self._lib.ccsp_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def launch_count() -> int:
    return int(load_library().ccsp_launch_count())


def reset_launch_count():
    load_library().ccsp_reset_launch_count()
