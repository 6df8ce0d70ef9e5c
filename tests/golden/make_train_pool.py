#!/usr/bin/env python
"""Build-container only: a TRAINING pool of RandomSplitQualitativeWorld scenes made with the reference's own scene generator and
labeller (the same recipe as tests/golden/make_scenes.py: envs/builders.py:10-52 + envs/data_utils.py:427-621, 408-415 through
oracle/ref_shim.py), stored compactly under diffusion_ccsp_b200/data/ so that the repo can train its own checkpoint anywhere.

    python tests/golden/make_train_pool.py [--scenes 24000]

Scene sizes are mixed (2..8 tiles; the reference trains on mixed sizes too, train_utils.py:153).  Seeds differ from the test
fixtures (tests/golden/scenes_qualitative_n*.npz), so the pools are disjoint draws.
"""
import argparse
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import make_scenes as ms  # noqa: E402
from oracle.ref_shim import load_reference, load_reference_envs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--scenes', type=int, default=24000)
    ap.add_argument('--out', default=os.path.join(ROOT, 'diffusion_ccsp_b200', 'data', 'scenes_qualitative_train.npz'))
    a = ap.parse_args()
    dfn, _ = load_reference()
    builders, du = load_reference_envs()
    qual = dfn.qualitative_constraints
    np.random.seed(20240); random.seed(20240)
    sizes = np.random.choice([2, 3, 4, 5, 6, 7, 8], size=a.scenes, p=[0.08, 0.10, 0.12, 0.14, 0.16, 0.18, 0.22])
    xs, eis, eas, ncount, ecount = [], [], [], [], []
    for n_obj in sizes:
        x, ei, ea, _ = ms.one_scene(builders, du, qual, int(n_obj))
        xs.append(x); eis.append(ei.astype(np.int8)); eas.append(ea.astype(np.int8))
        ncount.append(x.shape[0]); ecount.append(ea.shape[0])
    np.savez_compressed(a.out, x=np.concatenate(xs), edge_local=np.concatenate(eis, 1), edge_attr=np.concatenate(eas),
                        nodes_per_scene=np.array(ncount, np.int16), edges_per_scene=np.array(ecount, np.int16))
    print(f'{a.scenes} scenes, {sum(ncount)} nodes, {sum(ecount)} edges -> {a.out} ({os.path.getsize(a.out) / 1e6:.2f} MB)')


if __name__ == '__main__':
    main()
