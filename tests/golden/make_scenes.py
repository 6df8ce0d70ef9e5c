"""Generate RandomSplitQualitativeWorld scene-graph fixtures with the REFERENCE's own code.

Runs only in the build container (needs /root/reference).  Uses, unmodified:
  envs/builders.py:10-52    get_tray_splitting_gen          (region layouts)
  envs/data_utils.py:427-621 compute_qualitative_constraints (13-type labelling)
  envs/data_utils.py:408-415 randomize_unordered_constraints (random flips of symmetric relations)
and restates the few trimesh-dependent lines that cannot be imported here:
  envs/mesh_utils.py:174-191 create_tray (extents/centres of bottom + 4 walls, t=0.1, h=0.01)
  envs/mesh_utils.py:227-258 regions_to_meshes (shrink each region by ps ~ U(0.2*p, 0.2)^4, p=0)
  envs/worlds.py:279-286 + networks/data_transforms.py:101-109  node rows
      [w/W, l/L, x/(W/2), y/(L/2), cs, sn], yaw = pi/2 (or pi with w,l swapped when l > w)

Output: diffusion_ccsp_b200/data/scenes_qualitative_n{N}.npz  (x f32 [n,6], edge_index i32 [2,E],
edge_attr i8 [E], mask i8 [n]).     python tests/golden/make_scenes.py
"""
import math
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_shim import load_reference, load_reference_envs  # noqa: E402

W, L, H = 3.0, 2.0, 0.5


def tray_objects(w, l, h=0.01, t=0.1):
    return {
        'bottom': {'extents': (w, l, t), 'center': (0, 0, -t / 2)},
        'north': {'extents': (w, t, h), 'center': (0, (l + t) / 2, h / 2)},
        'south': {'extents': (w, t, h), 'center': (0, -(l + t) / 2, h / 2)},
        'west': {'extents': (t, l + 2 * t, h), 'center': (-(w + t) / 2, 0, h / 2)},
        'east': {'extents': (t, l + 2 * t, h), 'center': ((w + t) / 2, 0, h / 2)},
    }


def one_scene(builders, data_utils, qual, n_obj):
    max_depth = math.ceil(math.log2(n_obj)) + 1
    while True:
        gen = builders.get_tray_splitting_gen(num_samples=2, min_num_regions=n_obj,
                                              max_num_regions=n_obj, max_depth=max_depth)
        regions = next(gen(W, L))
        tiles = []
        for (x, y, w, l) in regions:
            ps = np.random.uniform(0.2 * 0.0, 0.2, 4)
            if w <= ps[1] + ps[3] or l <= ps[0] + ps[2]:
                continue
            w -= ps[1] + ps[3]; x += ps[1]; l -= ps[0] + ps[2]; y += ps[0]
            tiles.append(((w, l, H), (-W / 2 + x + w / 2, -L / 2 + y + l / 2, H / 2)))
        if len(tiles) == n_obj:
            break
    objects = tray_objects(W, L)
    for i, (ext, cen) in enumerate(tiles):
        objects[f'tile_{i}'] = {'extents': ext, 'center': cen}
    cons = [('in', i, 0) for i in range(1, n_obj + 1)]
    cons += [('cfree', i, j) for i in range(1, n_obj + 1) for j in range(i + 1, n_obj + 1)]
    q = data_utils.compute_qualitative_constraints(objects, rotations={}, scale=min(W / 3, L / 2))
    cons += data_utils.randomize_unordered_constraints(q)

    rows = [[1.0, 1.0, 0.0, 0.0, 0.0, 0.0]]
    for (w, l, _), (cx, cy, _) in tiles:
        yaw = math.pi / 2
        if l > w:
            w, l = l, w
            yaw = math.pi
        rows.append([w / W, l / L, cx / (W / 2), cy / (L / 2), math.cos(yaw), math.sin(yaw)])
    ei = np.array([[c[1], c[2]] for c in cons], dtype=np.int32).T
    ea = np.array([qual.index(c[0]) for c in cons], dtype=np.int8)
    mask = np.zeros(n_obj + 1, np.int8); mask[0] = 1
    return np.array(rows, np.float32), ei, ea, mask


def main():
    dfn, _ = load_reference()
    builders, data_utils = load_reference_envs()
    qual = dfn.qualitative_constraints
    for n_obj, count in ((4, 64), (8, 1024), (3, 16), (6, 16)):
        np.random.seed(n_obj); random.seed(n_obj)
        xs, eis, eas, ms, off = [], [], [], [], 0
        for _ in range(count):
            x, ei, ea, m = one_scene(builders, data_utils, qual, n_obj)
            xs.append(x); eis.append(ei + off); eas.append(ea); ms.append(m)
            off += x.shape[0]
        out = os.path.join(os.path.dirname(os.path.dirname(HERE)), 'diffusion_ccsp_b200', 'data', f'scenes_qualitative_n{n_obj}.npz')
        np.savez_compressed(out, x=np.concatenate(xs), edge_index=np.concatenate(eis, 1),
                            edge_attr=np.concatenate(eas), mask=np.concatenate(ms))
        E = sum(e.shape[0] for e in eas)
        hist = np.bincount(np.concatenate(eas), minlength=len(qual)) / count
        print(f'N={n_obj}: {count} scenes, {off} nodes, {E} edges ({E / count:.1f}/scene) -> {out} '
              f'({os.path.getsize(out) / 1024:.0f} KiB)')
        print('   per-scene type histogram:', dict(zip(qual, np.round(hist, 1))))


if __name__ == '__main__':
    main()
