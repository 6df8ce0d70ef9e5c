"""On-disk formats and the batch builder around the sampling path — host-side Python, like the reference.

  load_raw_graph / save_raw_graph      data/<name>/raw/data_i.pt            (envs/data_utils.py:118-127 save_graph_data)
  data_transform_cn_diffuse_batch      raw graph -> normalised scene        (networks/data_transforms.py:26-200)
  GraphDataset                         data/<name>/raw/*.pt -> scenes       (datasets.py:27-117)
  get_args_from_run_id                 wandb/<run>/files/config.yaml        (train_utils.py:316-337, flag defaults :86-111)
  load_checkpoint                      logs/<run>/model-<k>.pt              (networks/ddpm.py:496-514)

torch_geometric is not a dependency: a raw `data_i.pt` written by the reference is a pickled `torch_geometric.data.Data`;
it is read here with a stand-in unpickler that materialises any `torch_geometric.*` class as a plain attribute bag (both the
PyG 1.x layout, attributes in `__dict__`, and the PyG 2.x layout, `_store._mapping`, are understood).
"""
from __future__ import annotations

import os
import pickle
import types
from argparse import Namespace
from os.path import isdir, isfile, join
from typing import Dict, List, Optional

import numpy as np
import torch

from .scenes import (SceneBatch, puzzle_constraints, qualitative_constraints, robot_constraints, stability_constraints)

robot_qualitative_constraints = robot_constraints + qualitative_constraints


# =====================================================================================================================
# raw graphs
# =====================================================================================================================
class RawGraph:
    """What world.generate_pt writes (envs/worlds.py:247-358): x [n, 1 + F'] (first column: 0 container / 1 tile, then
    UN-normalised geometry and pose), edge_index = list of (constraint name, i, j), y = labels."""

    def __init__(self, x, edge_index, y=None):
        self.x = torch.as_tensor(x, dtype=torch.float32)
        self.edge_index = edge_index
        self.y = y


class _Bag:
    """stand-in for any torch_geometric class while unpickling"""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, tuple) and len(state) == 2 and isinstance(state[1], dict):   # (dict, slots)
            state = {**(state[0] or {}), **state[1]}
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__['_state'] = state


def _bag_class(mod, name):
    return type(name, (_Bag,), {'__module__': mod})


class _StubUnpickler(pickle.Unpickler):
    def find_class(self, mod, name):
        if mod.split('.')[0] == 'torch_geometric':
            try:
                return super().find_class(mod, name)          # the real class when PyG is installed
            except Exception:
                return _bag_class(mod, name)
        return super().find_class(mod, name)


_stub_pickle = types.SimpleNamespace(Unpickler=_StubUnpickler, load=lambda f, **k: _StubUnpickler(f, **k).load(),
                                     loads=pickle.loads, dump=pickle.dump, dumps=pickle.dumps, __name__='pickle',
                                     HIGHEST_PROTOCOL=pickle.HIGHEST_PROTOCOL, DEFAULT_PROTOCOL=pickle.DEFAULT_PROTOCOL)


def _field(obj, key):
    if isinstance(obj, dict):
        return obj.get(key)
    d = getattr(obj, '__dict__', {})
    if key in d:
        return d[key]
    store = d.get('_store')
    if store is not None:
        mapping = getattr(store, '__dict__', {}).get('_mapping')
        if isinstance(mapping, dict) and key in mapping:
            return mapping[key]
    try:
        return getattr(obj, key)
    except Exception:
        return None


def load_raw_graph(path: str) -> RawGraph:
    """Read `data_i.pt` (a pickled PyG Data written by the reference, or the plain dict `save_raw_graph` writes)."""
    obj = torch.load(path, map_location='cpu', pickle_module=_stub_pickle, weights_only=False)
    x, ei, y = _field(obj, 'x'), _field(obj, 'edge_index'), _field(obj, 'y')
    if x is None or ei is None:
        raise ValueError(f'{path}: no x / edge_index found')
    if torch.is_tensor(ei):                                   # plain index tensors [2, E] (non-named edges)
        ei = [tuple(int(v) for v in col) for col in ei.t().tolist()]
    return RawGraph(x, [tuple(e) for e in ei], y)


def save_raw_graph(path: str, nodes, edge_index, labels=None):
    """envs/data_utils.py:118-127 without PyG: the same three fields in a plain dict."""
    if labels is None or labels[0] is None:
        labels = [0] * len(nodes)
    torch.save(dict(x=torch.tensor(np.asarray(nodes), dtype=torch.float), edge_index=[tuple(e) for e in edge_index],
                    y=torch.tensor(np.asarray(labels), dtype=torch.float)), path)


# =====================================================================================================================
# networks/data_transforms.py:26-200
# =====================================================================================================================
def data_transform_cn_diffuse_batch(data: RawGraph, data_idx: int, input_mode: str) -> SceneBatch:
    """Normalise one raw graph into the sampler's input rows `[geom, pose]` (x / edge_index / edge_attr / mask / x_extract /
    edge_extract / world_dims).  Arithmetic is Python floats then float32, like the reference (`dd = data.x[i].tolist()`,
    :54; torch.tensor(..., dtype=torch.float), :187)."""
    features = []
    all_constraints = puzzle_constraints
    w_tray, l_tray = (float(v) for v in data.x[0, 1:3])
    world_dims = (w_tray, l_tray)
    for i in range(len(data.x)):
        dd = data.x[i].tolist()
        if len(dd) == 5:                                                             # :56-64 boxes
            typ, w, l, x, y = dd
            w /= w_tray; l /= l_tray; x /= (w_tray / 2); y /= (l_tray / 2)
            geom, pose = [w, l], [x, y]
        elif len(dd) == 7:
            if 'diffuse_pairwise' in input_mode:                                     # :68-86 triangle P1 encoding with theta
                if dd[0] == 0:
                    typ, w, l, _, x, y, _ = dd
                    w /= w_tray; l /= l_tray
                    geom, pose = [w, l, 0], [x, y, 0]
                else:
                    typ, l, x3, y3, x1, y1, r1 = dd
                    l /= w_tray; x3 /= w_tray; y3 /= l_tray; x1 /= (w_tray / 2); y1 /= (l_tray / 2); r1 /= np.pi
                    geom, pose = [l, x3, y3], [x1, y1, r1]
            else:                                                                    # :89-109 box encoding with sin/cos
                if dd[0] == 0:
                    typ, w, l, x, y, _, _ = dd
                    w /= w_tray; l /= l_tray
                    geom, pose = [w, l], [x, y, 0, 0]
                elif 'stability' in input_mode:
                    all_constraints = stability_constraints
                    geom, pose = dd[1:3], dd[3:]
                elif 'qualitative' in input_mode:
                    all_constraints = qualitative_constraints
                    _, w, l, x, y, sn, cs = dd
                    w /= w_tray; l /= l_tray; x /= (w_tray / 2); y /= (l_tray / 2)
                    geom, pose = [w, l], [x, y, cs, sn]
                else:
                    raise ValueError(f'7-column rows need a diffuse_pairwise / stability / qualitative input_mode, got {input_mode!r}')
        elif len(dd) == 8:                                                           # :112-130 triangle P1 with sin/cos
            if dd[0] == 0:
                typ, w, l, _, x, y, _, _ = dd
                w /= w_tray; l /= l_tray
                geom, pose = [w, l, 0], [x, y, 0, 0]
            else:
                typ, l, x3, y3, x1, y1, cs, sn = dd
                l /= w_tray; x3 /= w_tray; y3 /= l_tray; x1 /= (w_tray / 2); y1 /= (l_tray / 2)
                geom, pose = [l, x3, y3], [x1, y1, cs, sn]
        elif len(dd) in (22, 29, 36):                                                # :159-166 robot rows: as stored
            geom, pose = dd[1:9], dd[9:]
            all_constraints = robot_constraints
            world_dims = tuple(geom[3:5])
            if 'robot' in input_mode and 'qualitative' in input_mode:
                all_constraints = robot_qualitative_constraints
        else:
            raise NotImplementedError(f'{len(dd)}-column raw rows (image / centroid encodings) are out of scope')
        features.append(geom + pose)

    edge_attr = [all_constraints.index(e[0]) for e in data.edge_index]              # :174
    edge_index = [list(e[1:]) for e in data.edge_index]                             # :175
    x = torch.tensor(np.stack([np.asarray(f) for f in features]), dtype=torch.float)
    mask = torch.zeros(x.shape[0], dtype=torch.int8)
    mask[0] = 1                                                                      # conditioned_variables = [0]   (:46, 182-183)
    ei = torch.tensor(np.asarray(edge_index, dtype=np.int64).reshape(-1, 2), dtype=torch.int64).T
    return SceneBatch(x, ei, torch.tensor(edge_attr, dtype=torch.float), mask,
                      x_extract=torch.ones(x.shape[0]) * data_idx, edge_extract=torch.ones(ei.shape[1]) * data_idx,
                      world_dims=[world_dims])


def pre_transform(data, data_idx, input_mode, **kwargs):
    """networks/data_transforms.py:15-21 (the constraint-network branch)."""
    if 'diffuse_pairwise' in input_mode or 'robot' in input_mode or 'stability' in input_mode or 'qualitative' in input_mode:
        return [data_transform_cn_diffuse_batch(data, data_idx, input_mode)]
    raise NotImplementedError(f'input_mode={input_mode!r}: only the constraint-network transform is on this path')


# =====================================================================================================================
# datasets.py:27-117
# =====================================================================================================================
class GraphDataset:
    """`GraphDataset(dir_name, input_mode)`: data/<dir_name>/raw/data_{i}.pt for i < N, where N is the number in the directory
    name's parentheses (datasets.py:45), each run through the pre_transform; indexable list of single-scene SceneBatch objects.
    JSON-based datasets (robot / stability, datasets.py:47, 92-104) need the simulator-side converters and are not read."""

    def __init__(self, dir_name: str, input_mode: str = 'qualitative', root: str = 'data', pre_filter=None):
        if 'robot' in input_mode or 'stability' in input_mode:
            raise NotImplementedError('solution.json datasets (robot / stability) are produced by the PyBullet pipeline')
        self.dir_name, self.input_mode = dir_name, input_mode
        self.root = join(root, dir_name)
        self.length = int(eval(dir_name[dir_name.index('(') + 1:dir_name.index(')')], {'__builtins__': {}}))
        if 'object_i=' in dir_name:
            self.length = 1
        self.scenes: List[SceneBatch] = []
        for idx in range(self.length):
            raw = load_raw_graph(join(self.root, 'raw', f'data_{idx}.pt'))
            if pre_filter is not None and not pre_filter(raw):
                continue
            self.scenes.extend(pre_transform(raw, idx, input_mode))

    def __len__(self):
        return len(self.scenes)

    def __getitem__(self, i):
        return self.scenes[i]

    def __iter__(self):
        return iter(self.scenes)


# =====================================================================================================================
# train_utils.py:86-111, 316-337
# =====================================================================================================================
ARG_DEFAULTS = dict(timesteps=1000, model='Diffusion-CCSP', EBM=False, energy_wrapper=False, samples_per_step=10,
                    step_sizes='2*self.betas', train_task='None', train_small=1, train_name='', train_proj='correct_norm',
                    train_num_steps=300000, input_mode=None, ebm_per_steps=1, ev='ff', hidden_dim=256, normalize=True,
                    pretrained=False, use_wandb=False, run_id=None, test_tasks=None)


def get_args_from_run_id(run_id: str, wandb_roots=('wandb', 'wandb2')) -> Namespace:
    """Flags of a training run recovered from wandb/<...run_id...>/files/config.yaml: every `key: {value: v}` entry except
    train_batch_size / train_lr overrides the default, then the reference's per-run patches (train_utils.py:329-336)."""
    import yaml
    args = Namespace(**ARG_DEFAULTS)
    args.run_id = run_id
    found = None
    for root in wandb_roots:
        if isdir(root):
            hits = sorted(join(root, f) for f in os.listdir(root) if run_id in f)
            if hits:
                found = hits[0]
                break
    if found is None:
        raise FileNotFoundError(f'no wandb run directory containing {run_id!r} under {wandb_roots}')
    config = yaml.load(open(join(found, 'files', 'config.yaml'), 'r'), Loader=yaml.FullLoader)
    for k, v in config.items():
        if k in ('train_batch_size', 'train_lr'):
            continue
        if type(v) is dict:
            setattr(args, k, v['value'])
    if run_id == 'j8lenp74':
        args.pretrained = True
    if run_id in ('bo02mwbw', '4xt8u4n7', 'qi3dqq2l'):
        args.normalize = False
    if run_id in ('9xhbwmi9', 'ta4tsbz6'):
        args.energy_wrapper = True
    if run_id in ('ql30000e', 'jn49b39m', 'g38uz4uk', 'uyq4fd3u', 'oamtpoae', '6jrpn5vf'):
        args.model = 'StructDiffusion'
    if args.EBM == 'False':
        args.EBM = False
    return args


def load_checkpoint(path: str) -> Dict:
    """logs/<run>/model-<milestone>.pt -> {'step', 'model': GaussianDiffusion state_dict[, 'ema']}  (ddpm.py:496-514)."""
    data = torch.load(path, map_location='cpu', weights_only=False)
    if 'model' not in data:
        raise ValueError(f'{path}: not a Trainer checkpoint (no "model" entry)')
    return data
