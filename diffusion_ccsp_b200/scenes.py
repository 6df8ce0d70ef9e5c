"""Scene-graph batches: the input contract of the sampling path.

A `SceneBatch` is the duck-typed stand-in for the PyG `Batch` the reference hands to
`GaussianDiffusion.sample` (networks/data_transforms.py:181-200 + PyG collation): a disjoint
union of scene graphs, node 0 of every scene being the container (mask = 1).

    x           f32 [n, F]   rows = [geom, (extras), pose, (extras)]
    edge_index  i64 [2, E]   row 0 = arg1, row 1 = arg2, node ids offset per scene
    edge_attr   f32 [E]      constraint-type id (float, as in data_transforms.py:174,190)
    mask        i8  [n]      1 = pinned (container) node
    x_extract   f32 [n]      scene id per node;  edge_extract f32 [E] scene id per edge

Everything here is host-side (CPU tensors), exactly like the reference where the batch
stays on the CPU.  The generators below synthesise inputs of the five
BASELINE.json configs; they restate the *shape* of the reference's data (edge structure,
feature layout) and are not on the hot path.
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Sequence

import numpy as np
import torch

# constraint vocabularies (networks/denoise_fn.py:16-25)
puzzle_constraints = ['in', 'cfree']
robot_constraints = ['gin', 'gfree']
stability_constraints = ['within', 'supportedby', 'cfree']
qualitative_constraints = [
    'in', 'center-in', 'left-in', 'right-in', 'top-in', 'bottom-in',
    'cfree', 'left-of', 'top-of',
    'close-to', 'away-from', 'h-aligned', 'v-aligned'
]


class SceneBatch:
    """Plain container with the attributes the sampler reads (x, edge_index, edge_attr, mask)."""

    def __init__(self, x, edge_index, edge_attr, mask, x_extract=None, edge_extract=None,
                 world_dims=None):
        self.x = torch.as_tensor(x, dtype=torch.float32)
        self.edge_index = torch.as_tensor(edge_index, dtype=torch.int64)
        self.edge_attr = torch.as_tensor(edge_attr, dtype=torch.float32)
        self.mask = torch.as_tensor(mask, dtype=torch.int8)
        n, E = self.x.shape[0], self.edge_index.shape[1]
        if x_extract is None:
            # scenes are delimited by masked (container) rows
            x_extract = torch.cumsum(self.mask.to(torch.int64), 0) - 1
        self.x_extract = torch.as_tensor(x_extract, dtype=torch.float32)
        if edge_extract is None:
            edge_extract = self.x_extract[self.edge_index[0]] if E else torch.zeros(0)
        self.edge_extract = torch.as_tensor(edge_extract, dtype=torch.float32)
        self.world_dims = world_dims
        assert self.edge_index.shape[0] == 2 and self.edge_attr.shape[0] == E
        assert self.mask.shape[0] == n and self.x_extract.shape[0] == n

    # -- PyG-like conveniences -------------------------------------------------------
    @property
    def num_nodes(self) -> int:
        return int(self.x.shape[0])

    @property
    def num_edges(self) -> int:
        return int(self.edge_index.shape[1])

    @property
    def num_graphs(self) -> int:
        """scenes in the batch; ids are 0..S-1 after `collate` (a single dataset item may carry its dataset index instead,
        data_transforms.py:197-198)"""
        if not self.num_nodes:
            return 0
        lo, hi = int(self.x_extract.min().item()), int(self.x_extract.max().item())
        return hi - lo + 1

    def clone(self) -> "SceneBatch":
        return SceneBatch(self.x.clone(), self.edge_index.clone(), self.edge_attr.clone(),
                          self.mask.clone(), self.x_extract.clone(), self.edge_extract.clone(),
                          self.world_dims)

    def to(self, *a, **k):  # the reference never moves the batch; keep it on the host
        return self

    # -- scene-level slicing (used for sharding across ranks) ----------
    def scene_node_ranges(self) -> np.ndarray:
        """[num_graphs + 1] node offsets (scenes are contiguous in a collated batch)."""
        sid = self.x_extract.to(torch.int64).numpy()
        assert np.all(np.diff(sid) >= 0), "scenes must be contiguous"
        counts = np.bincount(sid, minlength=self.num_graphs)
        return np.concatenate([[0], np.cumsum(counts)])

    def select_scenes(self, lo: int, hi: int) -> "SceneBatch":
        """Sub-batch of scenes [lo, hi) with node ids re-based to start at 0."""
        off = self.scene_node_ranges()
        n0, n1 = int(off[lo]), int(off[hi])
        esid = self.edge_extract.to(torch.int64)
        keep = (esid >= lo) & (esid < hi)
        wd = self.world_dims[lo:hi] if isinstance(self.world_dims, list) and len(self.world_dims) == self.num_graphs else self.world_dims
        return SceneBatch(self.x[n0:n1].clone(), self.edge_index[:, keep] - n0,
                          self.edge_attr[keep].clone(), self.mask[n0:n1].clone(),
                          self.x_extract[n0:n1] - lo, self.edge_extract[keep] - lo, wd)

    def shard(self, rank: int, world_size: int) -> "SceneBatch":
        """Contiguous, near-equal scene shard for `rank` (no scene is split)."""
        lo, hi = shard_bounds(self.num_graphs, rank, world_size)
        return self.select_scenes(lo, hi)


def shard_bounds(num_scenes: int, rank: int, world_size: int):
    base, rem = divmod(num_scenes, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def collate(scenes: Sequence[SceneBatch]) -> SceneBatch:
    """PyG-style collation: concatenate nodes, offset edge ids per scene."""
    xs, eis, eas, ms, xe, ee = [], [], [], [], [], []
    off = 0
    sid = 0
    wd = []
    for s in scenes:
        g = s.num_graphs
        base = s.x_extract.min() if s.num_nodes else 0          # dataset items carry their dataset index: re-base to 0..S-1
        xs.append(s.x); eis.append(s.edge_index + off); eas.append(s.edge_attr); ms.append(s.mask)
        xe.append(s.x_extract - base + sid); ee.append(s.edge_extract - base + sid)
        off += s.num_nodes
        sid += g
        wd.append(s.world_dims)
    world_dims = None
    if all(w is not None for w in wd):                           # PyG collation concatenates the per-item world_dims lists
        world_dims = [tuple(d) for w in wd for d in w]
    return SceneBatch(torch.cat(xs), torch.cat(eis, 1), torch.cat(eas), torch.cat(ms),
                      torch.cat(xe), torch.cat(ee), world_dims)


def take_scenes(batch: SceneBatch, ids: Sequence[int]) -> SceneBatch:
    return collate([batch.select_scenes(int(i), int(i) + 1) for i in ids])


class SceneLoader:
    """Minimal stand-in for `torch_geometric.loader.DataLoader(dataset, batch_size, shuffle)` (ddpm.py:443-449) over a pool of
    scenes: yields collated SceneBatch objects of `batch_size` scenes (the last one may be smaller), reshuffled every epoch."""

    def __init__(self, pool, batch_size: int, shuffle: bool = False, seed: int = 0):
        if not isinstance(pool, SceneBatch):
            pool = collate(list(pool))
        self.pool = pool
        self.batch_size, self.shuffle = int(batch_size), shuffle
        self._rng = np.random.default_rng(seed)

    def __len__(self):
        return (self.pool.num_graphs + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        G = self.pool.num_graphs
        order = self._rng.permutation(G) if self.shuffle else np.arange(G)
        for i in range(0, G, self.batch_size):
            yield _gather_scenes_fast(self.pool, order[i:i + self.batch_size])


# =====================================================================================
# synthetic generators
# =====================================================================================

def _split_tray(rng: np.random.Generator, w: float, l: float, n_regions: int,
                min_frac: float = 0.4):
    """Random guillotine split of a w x l tray into exactly `n_regions` regions (same family of layouts as
    envs/builders.py:10-52).  The reference draws a random recursive partition and rejects until the count
    matches, which takes minutes for 12 regions; here the largest region is split (uniform cut in the middle
    40 % of its longer side) until the count is reached, so every draw succeeds."""
    regs = [(0.0, 0.0, w, l)]
    while len(regs) < n_regions:
        k = int(np.argmax([r[2] * r[3] for r in regs]))
        x, y, bw, bl = regs.pop(k)
        f = rng.uniform(0.3, 0.7)
        if bw >= bl:
            regs += [(x, y, bw * f, bl), (x + bw * f, y, bw * (1 - f), bl)]
        else:
            regs += [(x, y, bw, bl * f), (x, y + bl * f, bw, bl * (1 - f))]
    order = rng.permutation(len(regs))
    return [regs[i] for i in order]


def _edges_in_cfree(n_obj: int):
    """('in', i, 0) for all i, ('cfree', i, j) for i<j   (envs/worlds.py:138-144)."""
    e = [(0, i, 0) for i in range(1, n_obj + 1)]
    e += [(1, i, j) for i in range(1, n_obj + 1) for j in range(i + 1, n_obj + 1)]
    return e


def boxes_scene(rng: np.random.Generator, n_obj: int, W: float = 3.0, L: float = 2.0) -> SceneBatch:
    """RandomSplitWorld scene, `diffuse_pairwise` rows [w/W, l/L, x/(W/2), y/(L/2)]  (F=4, P=2)."""
    while True:
        regs = _split_tray(rng, W, L, n_obj)
        rows = [[1.0, 1.0, 0.0, 0.0]]
        for (x, y, w, l) in regs:
            ps = rng.uniform(0.02, 0.2, 4) * min(1.0, 2.0 * min(w, l))     # padding shrinks with the region
            if w <= ps[1] + ps[3] or l <= ps[0] + ps[2]:
                break
            w2, l2 = w - ps[1] - ps[3], l - ps[0] - ps[2]
            cx, cy = -W / 2 + x + ps[1] + w2 / 2, -L / 2 + y + ps[0] + l2 / 2
            rows.append([w2 / W, l2 / L, cx / (W / 2), cy / (L / 2)])
        if len(rows) == n_obj + 1:
            break
    e = np.array(_edges_in_cfree(n_obj), dtype=np.int64)
    mask = np.zeros(n_obj + 1, np.int8); mask[0] = 1
    return SceneBatch(np.array(rows, np.float32), e[:, 1:].T.copy(), e[:, 0].astype(np.float32), mask)


def triangles_scene(rng: np.random.Generator, n_obj: int) -> SceneBatch:
    """TriangularRandomSplitWorld scene, rows [l/W, x3/W, y3/L, x1/(W/2), y1/(L/2), cs, sn] (F=7, P=4)."""
    rows = [[1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0]]
    for _ in range(n_obj):
        th = rng.uniform(-math.pi, math.pi)
        rows.append([rng.uniform(0.2, 0.8), rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5),
                     rng.uniform(-0.9, 0.9), rng.uniform(-0.9, 0.9), math.cos(th), math.sin(th)])
    e = np.array(_edges_in_cfree(n_obj), dtype=np.int64)
    mask = np.zeros(n_obj + 1, np.int8); mask[0] = 1
    return SceneBatch(np.array(rows, np.float32), e[:, 1:].T.copy(), e[:, 0].astype(np.float32), mask)


def robot_box_scene(rng: np.random.Generator, n_obj: int) -> SceneBatch:
    """TableToBoxWorld scene, 28-column rows (networks/data_transforms.py:203-269):
    geom 0-7, id/scale 8-9, grasp one-hot 10-14, grasp_id 15, pose 16-20 [x,y,z,sn,cs], pick pose 21-27.
    Edges ('gin', i, 0) and ('gfree', j, i), j>i."""
    w0, l0, h0 = rng.uniform(0.3, 0.5), rng.uniform(0.3, 0.6), 0.25
    x0, y0 = rng.uniform(0.2, 0.4), rng.uniform(-0.1, 0.1)
    rows = [[1, 1, 1, w0, l0, h0, x0, y0] + [0] * 20]
    for i in range(n_obj):
        w, l, h = rng.uniform(0.1, 0.5), rng.uniform(0.1, 0.5), rng.uniform(0.2, 1.0)
        side = np.zeros(5); side[rng.integers(0, 5)] = 1
        th = rng.uniform(-math.pi, math.pi)
        pose = [rng.uniform(-0.9, 0.9), rng.uniform(-0.9, 0.9), h / 2, math.sin(th), math.cos(th)]
        pick = list(rng.uniform(-1, 1, 7))
        rows.append([w, l, h, w0, l0, h0, x0, y0, i + 1, 1.0] + list(side) + [float(rng.integers(0, 20))]
                    + pose + pick)
    e = [(0, i, 0) for i in range(1, n_obj + 1)]
    e += [(1, j, i) for i in range(1, n_obj + 1) for j in range(i + 1, n_obj + 1)]
    e = np.array(e, dtype=np.int64)
    mask = np.zeros(n_obj + 1, np.int8); mask[0] = 1
    return SceneBatch(np.array(rows, np.float32), e[:, 1:].T.copy(), e[:, 0].astype(np.float32), mask)


def random_typed_scene(rng: np.random.Generator, n_obj: int, n_types: int, n_edges: int,
                       F: int, allow_isolated: bool = False) -> SceneBatch:
    """Arbitrary typed multigraph for property tests (random types/endpoints, self-pairs excluded)."""
    x = rng.uniform(-1, 1, (n_obj + 1, F)).astype(np.float32)
    a = rng.integers(0, n_obj + 1, n_edges)
    b = (a + rng.integers(1, n_obj + 1, n_edges)) % (n_obj + 1)
    t = rng.integers(0, n_types, n_edges)
    if not allow_isolated:   # make sure every node has an incident edge (deg 0 => NaN, denoise_fn.py:524)
        extra_a = np.arange(n_obj + 1)
        extra_b = (extra_a + 1) % (n_obj + 1)
        a, b = np.concatenate([a, extra_a]), np.concatenate([b, extra_b])
        t = np.concatenate([t, rng.integers(0, n_types, n_obj + 1)])
    mask = np.zeros(n_obj + 1, np.int8); mask[0] = 1
    return SceneBatch(x, np.stack([a, b]).astype(np.int64), t.astype(np.float32), mask)


def make_batch(kind: str, num_scenes: int, n_obj: int, seed: int = 0) -> SceneBatch:
    """kind in {'boxes', 'triangles', 'robot_box'} — the structured configs (3, 4, 5)."""
    rng = np.random.default_rng(seed)
    fn = {'boxes': boxes_scene, 'triangles': triangles_scene, 'robot_box': robot_box_scene}[kind]
    return collate([fn(rng, n_obj) for _ in range(num_scenes)])


# -------------------------------------------------------------------------------------
# RandomSplitQualitativeWorld scenes: committed fixtures (diffusion_ccsp_b200/data/) generated with the reference's own
# scene generator + qualitative labeller (tests/golden/make_scenes.py, tests/golden/make_train_pool.py); tiled to any batch.
# -------------------------------------------------------------------------------------
_FIXTURE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')


def load_scene_fixture(path: str) -> SceneBatch:
    z = np.load(path)
    return SceneBatch(z['x'], z['edge_index'].astype(np.int64), z['edge_attr'].astype(np.float32), z['mask'])


def qualitative_batch(num_scenes: int, n_obj: int = 8, seed: int = 0,
                      fixture_dir: Optional[str] = None) -> SceneBatch:
    """`num_scenes` RandomSplitQualitativeWorld scenes with `n_obj` tiles.  Scenes are drawn from
    the committed fixture pool `diffusion_ccsp_b200/data/scenes_qualitative_n{n_obj}.npz`; when more scenes than
    the pool holds are requested the pool is re-sampled with a seeded permutation (scenes are
    independent, so repeats only matter for statistics, not for the work per scene)."""
    path = os.path.join(fixture_dir or _FIXTURE_DIR, f'scenes_qualitative_n{n_obj}.npz')
    pool = load_scene_fixture(path)
    G = pool.num_graphs
    if num_scenes <= G and seed == 0:
        return pool.select_scenes(0, num_scenes)
    rng = np.random.default_rng(seed)
    ids = np.concatenate([rng.permutation(G) for _ in range((num_scenes + G - 1) // G)])[:num_scenes]
    return _gather_scenes_fast(pool, ids)


def qualitative_train_pool(path: Optional[str] = None) -> SceneBatch:
    """The committed TRAINING pool (24 000 RandomSplitQualitativeWorld scenes with 2..8 tiles, tests/golden/make_train_pool.py):
    disjoint draws from the evaluation fixtures, stored with scene-local edge ids."""
    z = np.load(path or os.path.join(_FIXTURE_DIR, 'scenes_qualitative_train.npz'))
    ncount, ecount = z['nodes_per_scene'].astype(np.int64), z['edges_per_scene'].astype(np.int64)
    off = np.concatenate([[0], np.cumsum(ncount)])[:-1]
    ei = z['edge_local'].astype(np.int64) + np.repeat(off, ecount)[None, :]
    mask = np.zeros(int(ncount.sum()), np.int8)
    mask[off] = 1
    return SceneBatch(z['x'], ei, z['edge_attr'].astype(np.float32), mask)


def _gather_scenes_fast(pool: SceneBatch, ids: np.ndarray) -> SceneBatch:
    """Vectorised take_scenes for large batches (per-pool offset tables are computed once and kept on the pool)."""
    tab = getattr(pool, '_gather_tables', None)
    if tab is None:
        off = pool.scene_node_ranges()
        esid = pool.edge_extract.to(torch.int64).numpy()
        order = np.argsort(esid, kind='stable')
        eoff = np.concatenate([[0], np.cumsum(np.bincount(esid, minlength=pool.num_graphs))])
        ei = pool.edge_index.numpy()[:, order]
        ei_local = ei - off[esid[order]][None, :]                    # node ids relative to the scene's first node
        tab = pool._gather_tables = (off, eoff, pool.x.numpy(), ei_local, pool.edge_attr.numpy()[order], pool.mask.numpy())
    off, eoff, x, ei_local, ea, m = tab
    ids = np.asarray(ids, dtype=np.int64)
    ncount, ecount = off[ids + 1] - off[ids], eoff[ids + 1] - eoff[ids]
    new_off = np.concatenate([[0], np.cumsum(ncount)])
    nsel = np.concatenate([np.arange(off[s], off[s + 1]) for s in ids]) if ids.size else np.zeros(0, np.int64)
    esel = np.concatenate([np.arange(eoff[s], eoff[s + 1]) for s in ids]) if ids.size else np.zeros(0, np.int64)
    shift = np.repeat(new_off[:-1], ecount)
    sid_n = np.repeat(np.arange(ids.size), ncount).astype(np.float32)
    sid_e = np.repeat(np.arange(ids.size), ecount).astype(np.float32)
    wd = pool.world_dims
    if isinstance(wd, list) and len(wd) == pool.num_graphs:
        wd = [wd[int(s)] for s in ids]
    return SceneBatch(x[nsel], ei_local[:, esel] + shift[None, :], ea[esel], m[nsel], sid_n, sid_e, wd)
