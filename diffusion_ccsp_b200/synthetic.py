"""Deterministic synthetic checkpoints and noise (no datasets / checkpoints exist offline).

`make_state_dict` produces a `GaussianDiffusion.state_dict()`-shaped dict — the 12 schedule buffers
are NOT included, only the `denoise_fn.*` keys listed in SURVEY.md §8b — with nn.Linear-style
U(-1/sqrt(fan_in), 1/sqrt(fan_in)) values drawn from numpy's PCG64 (bit-stable across machines,
unlike relying on torch's default init order).  The golden generator loads the same dict into the
unmodified reference with `load_state_dict`, the tests load it into the CUDA path.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch

from .scenes import (puzzle_constraints, qualitative_constraints, robot_constraints,
                     stability_constraints)

DIMS = {  # train_utils.py:266-278
    'qualitative': ((2, 0, 2), (4, 2, 6)),
    'stability_flat': ((2, 0, 2), (4, 2, 6)),
    'diffuse_pairwise': ((2, 0, 2), (2, 2, 4)),
    'diffuse_pairwise_triangular': ((3, 0, 3), (4, 3, 7)),
    'robot_box': ((8, 0, 8), (5, 10, 15), (5, 16, 21)),
}


def dims_for(input_mode: str, triangular: bool = False):
    """Substring dispatch of train_utils.py:266-278 ('robot' / 'stability' / 'qualitative' in input_mode)."""
    if 'robot' in input_mode:
        return DIMS['robot_box']
    if 'stability' in input_mode:
        return DIMS['stability_flat']
    if 'qualitative' in input_mode:
        return DIMS['qualitative']
    if triangular:
        return DIMS['diffuse_pairwise_triangular']
    return DIMS['diffuse_pairwise']


def constraint_set_for(input_mode: str):
    """networks/denoise_fn.py:207-214."""
    if 'robot' in input_mode:
        return robot_constraints
    if 'stability' in input_mode:
        return stability_constraints
    if 'qualitative' in input_mode:
        return qualitative_constraints
    return puzzle_constraints


def linear_shapes(dims, input_mode: str, hidden_dim: int = 256) -> Dict[str, Tuple[int, int]]:
    """name -> (out_features, in_features) for every nn.Linear of ConstraintDiffuser
    (networks/denoise_fn.py:227-264, 293-308), in construction order."""
    H = hidden_dim
    G, P = dims[0][0], dims[-1][0]
    robot = 'robot' in input_mode
    shapes = {'geom_encoder.0': (H // 2, G), 'geom_encoder.2': (H, H // 2)}
    if robot:
        shapes.update({'grasp_encoder.0': (H // 2, dims[1][0]), 'grasp_encoder.2': (H, H // 2)})
    shapes.update({
        'pose_encoder.0': (H // 2, P), 'pose_encoder.2': (H, H // 2),
        'pose_decoder.0': (H // 2, H), 'pose_decoder.2': (P, H // 2),
        'time_mlp.1': (4 * H, H), 'time_mlp.3': (H, 4 * H),
    })
    k_in = H * (6 if robot else 5)
    for c in range(len(constraint_set_for(input_mode))):
        shapes[f'mlps.{c}.0'] = (2 * H, k_in)
    return shapes


def make_state_dict(dims, input_mode: str, hidden_dim: int = 256, seed: int = 0,
                    prefix: str = 'denoise_fn.') -> Dict[str, torch.Tensor]:
    rng = np.random.default_rng(seed)
    sd = {}
    for name, (o, i) in linear_shapes(dims, input_mode, hidden_dim).items():
        bound = 1.0 / np.sqrt(i)
        sd[f'{prefix}{name}.weight'] = torch.from_numpy(rng.uniform(-bound, bound, (o, i)).astype(np.float32))
        sd[f'{prefix}{name}.bias'] = torch.from_numpy(rng.uniform(-bound, bound, (o,)).astype(np.float32))
    return sd


def make_noise(T: int, K: int, n: int, P: int, seed: int = 123) -> torch.Tensor:
    """[1 + T(1+K), n, P] float32 standard normals: draw 0 is x_T, then per timestep one p_sample
    draw + K ULA draws (reference draw order)."""
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.standard_normal((1 + T * (1 + K), n, P), dtype=np.float32))


def make_trained_state_dict(fixture_path: str = None) -> Dict[str, torch.Tensor]:
    """Qualitative-mode state_dict in the REALISTIC regime: synthetic.make_state_dict(seed) with the ~74 k
    parameters trained by tests/golden/make_trained_fixture.py (pose encoder/decoder, mlps biases; trained with
    the reference's own loss) overlaid.  With it the sampler stays at |x| = O(1) like a real checkpoint."""
    import os
    if fixture_path is None:
        fixture_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'trained_small_qualitative.npz')
    z = np.load(fixture_path)
    sd = make_state_dict(DIMS['qualitative'], 'qualitative', seed=int(z['weight_seed']))
    for k in z.files:
        if k != 'weight_seed':
            assert k in sd and tuple(sd[k].shape) == z[k].shape, k
            sd[k] = torch.from_numpy(z[k].astype(np.float32))
    return sd


def load_trained_checkpoint(path: str = None) -> Dict[str, torch.Tensor]:
    """The qualitative-world checkpoint this repo trained ITSELF (scripts/train_fixture.py: GaussianDiffusion.forward ->
    ccsp_train_step + ccsp_adam_step on the committed 24k-scene pool, no reference code involved), stored in fp16
    (diffusion_ccsp_b200/data/denoise_fn_qualitative_fp16.npz, 9.15 M parameters) and widened to fp32 here — the fp32 values ARE
    the checkpoint, for this repo and for the reference alike.  Selected at 15 000 steps: every N = 8 scene samples finite and
    the N = 4 solved rate is 22 % per try (profiles/train_fixture_r2e.log)."""
    import os
    if path is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'denoise_fn_qualitative_fp16.npz')
    z = np.load(path)
    return {k: torch.from_numpy(z[k].astype(np.float32)) for k in z.files if k.startswith('denoise_fn.')}
