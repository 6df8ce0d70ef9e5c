"""SolvedChecker — the success check of Trainer.evaluate on the GPU.

Host-side mirror of the per-graph CPU loop of networks/ddpm.py:620-713 (`render_world_from_graph` ->
`world.check_constraints_satisfied`, envs/data_utils.py:221-357, envs/worlds.py:734-764, envs/collisions.py:58-130)
for the 2-D box worlds: the batch's scene structure is compiled once (scene CSR, per-scene edge offsets as at
ddpm.py:690-691, world dims) and every sampled pose tensor is then checked by ONE launch of `ccsp_check_solved`
(C ABI) — the poses never leave the device, the result is S bytes.

No CPU fallback: without the CUDA library this raises (oracle/checker_oracle.py is test infrastructure only).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _abi

WORLD_BOXES, WORLD_QUALITATIVE = 0, 1


def world_kind_for(input_mode: str) -> Optional[int]:
    """which worlds the kernel covers (the rest needs trimesh / FCL Convex / PyBullet, ddpm.py:650-668)"""
    if 'qualitative' in input_mode and 'robot' not in input_mode:
        return WORLD_QUALITATIVE
    if input_mode == 'diffuse_pairwise':
        return WORLD_BOXES
    return None


class CheckDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ('kind', 'num_scenes', 'F', 'P', 'pose_begin', 'clamp')] + \
               [(n, C.c_void_p) for n in ('x', 'scene_node_ptr', 'scene_edge_ptr', 'edge_a', 'edge_b', 'edge_type', 'world_dims')]


def scene_tables(batch, default_world_dims=(3, 2)):
    """Host-side compilation of the scene structure: (scene_node_ptr [S+1], scene_edge_ptr [S+1], edge_a, edge_b, edge_type,
    world_dims [S,2]); edges grouped by scene in their original order, endpoints re-based by the scene's smallest edge index
    (ddpm.py:688-691)."""
    sid = batch.x_extract.to(torch.int64).numpy()
    if sid.size and np.any(np.diff(sid) < 0):
        raise ValueError('scenes must be contiguous in the batch')
    S = int(sid.max()) + 1 if sid.size else 0
    node_ptr = np.concatenate([[0], np.cumsum(np.bincount(sid, minlength=S))]).astype(np.int32)
    esid = batch.edge_extract.to(torch.int64).numpy()
    ei = batch.edge_index.numpy()
    ea = batch.edge_attr.numpy()
    if ea.size and (ea.min() < 0 or np.any(ea != np.floor(ea))):
        raise ValueError('edge_attr must hold non-negative integer type ids')
    order = np.argsort(esid, kind='stable')
    edge_ptr = np.concatenate([[0], np.cumsum(np.bincount(esid, minlength=S))]).astype(np.int32)
    a, b = ei[0][order], ei[1][order]
    if order.size:
        off = np.zeros(S, np.int64)                      # scenes without edges have no reduceat segment
        nonempty = np.where(np.diff(edge_ptr) > 0)[0]
        off[nonempty] = np.minimum.reduceat(np.minimum(a, b), edge_ptr[nonempty])
        per_edge = np.repeat(off, np.diff(edge_ptr))
        a, b = a - per_edge, b - per_edge
    wd = getattr(batch, 'world_dims', None)
    if wd is None:
        wd = [default_world_dims] * S
    wd = np.asarray([tuple(w) for w in wd], dtype=np.float32).reshape(S, 2)
    return node_ptr, edge_ptr, a.astype(np.int32), b.astype(np.int32), ea[order].astype(np.int32), wd


class SolvedChecker:
    """checker = SolvedChecker(batch, dims, input_mode, device);  solved = checker(poses)  ->  bool tensor [S] on the device.

    `poses` is what `GaussianDiffusion.sample` returns ([n,P] on the device, unclamped: the clamp of ddpm.py:620 happens in
    the kernel)."""

    def __init__(self, batch, dims, input_mode: str, device, world_dims=(3, 2)):
        kind = world_kind_for(input_mode)
        if kind is None:
            raise NotImplementedError(f'input_mode={input_mode!r}: only the 2-D box worlds are checked on the GPU '
                                      '(triangles / 3-D / robot need trimesh, FCL Convex or PyBullet)')
        self._lib = _abi.load_library()
        if not torch.cuda.is_available():
            raise _abi.CcspError('CUDA device required: the solved-checker has no CPU fallback')
        self.device = torch.device(device)
        self.kind = kind
        self.F = int(batch.x.shape[1])
        self.P, self.pose_begin = int(dims[-1][0]), int(dims[-1][1])
        node_ptr, edge_ptr, a, b, typ, wd = scene_tables(batch, world_dims)
        if np.any(wd[:, 0] == 3) and np.any((wd[:, 0] == 3) & (wd[:, 1] == 3)) and self.F == 6:
            raise NotImplementedError('3 x 3 worlds with 6-feature rows are the triangle P1 encoding (data_utils.py:239-247)')
        if node_ptr.size > 1 and int(np.diff(node_ptr).max()) - 1 > 28:
            raise ValueError('at most 28 tiles per scene')
        self.num_scenes = int(node_ptr.size - 1)
        self.n = int(batch.x.shape[0])
        dev = self.device
        t = lambda arr: torch.from_numpy(np.ascontiguousarray(arr)).to(dev)
        self._x = batch.x.detach().to(dev, torch.float32).contiguous()
        self._node_ptr, self._edge_ptr = t(node_ptr), t(edge_ptr)
        self._a, self._b, self._typ, self._wd = t(a), t(b), t(typ), t(wd)
        self.h2d_bytes = sum(int(v.numel() * v.element_size()) for v in (self._x, self._node_ptr, self._edge_ptr, self._a, self._b, self._typ, self._wd))
        d = CheckDesc()
        d.kind, d.num_scenes, d.F, d.P, d.pose_begin, d.clamp = kind, self.num_scenes, self.F, self.P, self.pose_begin, 1
        d.x, d.scene_node_ptr, d.scene_edge_ptr = self._x.data_ptr(), self._node_ptr.data_ptr(), self._edge_ptr.data_ptr()
        d.edge_a, d.edge_b, d.edge_type = self._a.data_ptr(), self._b.data_ptr(), self._typ.data_ptr()
        d.world_dims = self._wd.data_ptr()
        self._desc = d
        self._lib.ccsp_check_solved.argtypes = [C.POINTER(CheckDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]

    def __call__(self, poses: torch.Tensor, clamp: bool = True, return_counts: bool = False):
        dev = self.device
        poses = poses.detach().to(dev, torch.float32).contiguous()
        if tuple(poses.shape) != (self.n, self.P):
            raise ValueError(f'poses must be [{self.n}, {self.P}], got {tuple(poses.shape)}')
        solved = torch.empty((self.num_scenes,), dtype=torch.uint8, device=dev)
        counts = torch.empty((self.num_scenes, 2), dtype=torch.int32, device=dev) if return_counts else None
        self._desc.clamp = int(bool(clamp))
        with torch.cuda.device(dev):
            _abi.check(self._lib.ccsp_check_solved(C.byref(self._desc), poses.data_ptr(), solved.data_ptr(),
                                                   counts.data_ptr() if counts is not None else None,
                                                   _abi.current_stream_ptr(dev)), 'ccsp_check_solved')
        solved = solved.bool()
        return (solved, counts) if return_counts else solved
