"""Shared helpers for the test-suite (golden loading, tolerances)."""
import os

import numpy as np

from diffusion_ccsp_b200 import scenes, synthetic

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def golden_names(prefix):
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith(prefix) and f.endswith('.npz'))


def load_golden(name):
    z = dict(np.load(os.path.join(GOLDEN, name + '.npz')))
    batch = None
    if 'edge_index' in z:
        batch = scenes.SceneBatch(z['x'], z['edge_index'].astype(np.int64),
                                  z['edge_attr'].astype(np.float32), z['mask'])
    return z, batch


def case_model(z):
    mode = str(z['input_mode'])
    dims = synthetic.dims_for(mode, bool(z['triangular']))
    sd = synthetic.make_state_dict(dims, mode, seed=int(z['weight_seed']))
    return mode, dims, sd


def rel_err(a, b):
    """max|a-b| / max(1, max|b|): the parity measure of SURVEY.md §8c."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b)))))


def rel_err_strict(a, b):
    """max|a-b| / max|b| — no floor at 1: what single denoiser evaluations (|out| ~ 0.3) are held to."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))
