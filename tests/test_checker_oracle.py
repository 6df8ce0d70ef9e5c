"""CPU: pin oracle/checker_oracle.py (the restatement of the reference's success check).

  * qualitative_relations == the reference's OWN compute_qualitative_constraints (envs/data_utils.py:427-621, imported unmodified
    through oracle/ref_shim.py) as a SET, on thousands of random, near-threshold and rotated layouts;
  * the 2-D SAT against hand-built touching / overlapping / rotated cases (FCL is absent: see the module header);
  * the whole per-graph check on the committed scene fixtures: ground-truth layouts are solved, perturbed ones are not.
"""
import math

import numpy as np
import pytest

from diffusion_ccsp_b200 import scenes
from oracle import checker_oracle as chk
from oracle import ref_shim

needs_ref = pytest.mark.skipif(not ref_shim.reference_available(), reason='reference sources not available')


def tray(w=3.0, l=2.0, h=0.01, t=0.1):
    return {
        'bottom': {'extents': (w, l, t), 'center': (0, 0, -t / 2)},
        'north': {'extents': (w, t, h), 'center': (0, (l + t) / 2, h / 2)},
        'south': {'extents': (w, t, h), 'center': (0, -(l + t) / 2, h / 2)},
        'west': {'extents': (t, l + 2 * t, h), 'center': (-(w + t) / 2, 0, h / 2)},
        'east': {'extents': (t, l + 2 * t, h), 'center': ((w + t) / 2, 0, h / 2)},
    }


def random_layout(rng, n, snap=False):
    objs = tray()
    rot = {}
    grid = [0.05, 0.1, 0.3, 0.5, 0.6, 0.0, -0.05]
    for i in range(n):
        w, l = rng.uniform(0.1, 1.2, 2)
        x, y = rng.uniform(-1.5, 1.5), rng.uniform(-1.0, 1.0)
        if snap and i > 0 and rng.random() < 0.7:
            # put this tile at a gap that sits on (or a hair either side of) a threshold relative to a previous tile
            j = rng.integers(0, i)
            o = objs[f'tile_box_{j}']
            (ox, oy, _), (ow, ol, _) = o['center'], o['extents']
            if f'tile_box_{j}' in rot and abs(abs(rot[f'tile_box_{j}']) - np.pi / 2) < 0.1:
                ow, ol = ol, ow
            d = float(rng.choice(grid)) + float(rng.choice([0.0, 1e-12, -1e-12, 1e-7, -1e-7]))
            if rng.random() < 0.5:
                x = ox + ow / 2 + d + w / 2
                y = oy + rng.choice([0.0, 0.04999999, 0.05, 0.3 * l])
            else:
                y = oy + ol / 2 + d + l / 2
                x = ox + rng.choice([0.0, 0.04999999, 0.05, 0.3 * w])
        objs[f'tile_box_{i}'] = {'extents': (float(w), float(l), 0.01), 'center': (float(x), float(y), 0.005)}
        r = rng.random()
        if r < 0.3:
            rot[f'tile_box_{i}'] = float(rng.choice([0.0, np.pi / 2, -np.pi / 2, np.pi / 2 - 0.0999, np.pi / 2 + 0.1001,
                                                     -np.pi / 2 + 0.09, np.pi, -3.0, 1.0]))
        elif r < 0.5:
            rot[f'tile_box_{i}'] = float(rng.uniform(-np.pi, np.pi))
    return objs, rot


@needs_ref
@pytest.mark.parametrize('snap', [False, True])
def test_relations_match_the_reference_labeller(snap):
    _, du = ref_shim.load_reference_envs()
    rng = np.random.default_rng(11 + snap)
    n_rel = 0
    for it in range(1500):
        n = int(rng.integers(1, 11))
        objs, rot = random_layout(rng, n, snap)
        rotations = rot if it % 3 else None
        scale = 1 if it % 5 else float(rng.uniform(0.5, 1.5))
        ref = du.compute_qualitative_constraints(objs, rotations=rotations, scale=scale)
        ref = set(chk.expand_unordered_constraints([tuple(c) for c in ref]))
        got = chk.qualitative_relations(objs, rotations=rotations, scale=scale)
        got = set(chk.expand_unordered_constraints(sorted(got)))
        assert got == ref, (it, sorted(got ^ ref))
        n_rel += len(ref)
    assert n_rel > 20000      # the comparison is not vacuous


@needs_ref
def test_relations_on_the_reference_generated_fixtures():
    """the committed scene fixtures were labelled by the reference; re-deriving with the port from the stored rows gives back
    every stored relation (stored = a random orientation of each symmetric relation)"""
    pool = scenes.qualitative_batch(64, 8)
    solved, ncol, nmiss = chk.check_batch(pool.x[:, 2:6].numpy(), pool, (2, 6))
    assert solved.all(), (ncol, nmiss)


def test_sat_hand_built_cases():
    c = chk.boxes_collide
    assert c((0, 0), (1, 1), None, (0.9, 0), (1, 1), None)                  # overlapping
    assert not c((0, 0), (1, 1), None, (1.5, 0), (1, 1), None)              # apart
    assert c((0, 0), (1, 1), None, (1.0, 0), (1, 1), None)                  # touching faces: FCL reports a collision (s > 0 separates)
    assert c((0, 0), (1, 1), 0.0, (1.0, 1.0), (1, 1), 0.0)                  # touching corners
    assert not c((0, 0), (1, 1), None, (1.0 + 1e-12, 0), (1, 1), None)
    # a unit square rotated by 45 deg reaches sqrt(2)/2 along x
    r = math.sqrt(2) / 2
    assert c((0, 0), (1, 1), math.pi / 4, (0.5 + r - 1e-9, 0), (1, 1), None)
    assert not c((0, 0), (1, 1), math.pi / 4, (0.5 + r + 1e-9, 0), (1, 1), None)
    # separated only along an axis of the ROTATED box (the axis-aligned projections overlap)
    assert not c((0, 0), (2.0, 0.2), math.pi / 4, (0.9, -0.2), (0.4, 0.4), None)
    assert c((0, 0), (2.0, 0.2), math.pi / 4, (0.6, 0.5), (0.4, 0.4), None)
    # +yaw with the UNROTATED extents (collisions.py:108-111): a 2 x 0.2 bar turned by pi/2 is tall, not wide
    assert not c((0, 0), (2.0, 0.2), math.pi / 2, (0.5, 0), (0.4, 0.4), None)
    assert c((0, 0), (2.0, 0.2), math.pi / 2, (0, 0.9), (0.4, 0.4), None)
    # symmetry
    rng = np.random.default_rng(0)
    for _ in range(2000):
        a = (rng.uniform(-1, 1), rng.uniform(-1, 1)); b = (rng.uniform(-1, 1), rng.uniform(-1, 1))
        ea, eb = rng.uniform(0.1, 1, 2), rng.uniform(0.1, 1, 2)
        ya, yb = rng.uniform(-3.2, 3.2), rng.uniform(-3.2, 3.2)
        assert c(a, ea, ya, b, eb, yb) == c(b, eb, yb, a, ea, ya)


def test_sat_against_polygon_clipping():
    """independent check of the SAT: two convex polygons overlap iff Sutherland-Hodgman clipping leaves a non-empty area"""
    def corners(c, e, yaw):
        cs, sn = math.cos(yaw), math.sin(yaw)
        pts = [(-e[0] / 2, -e[1] / 2), (e[0] / 2, -e[1] / 2), (e[0] / 2, e[1] / 2), (-e[0] / 2, e[1] / 2)]
        return [(c[0] + cs * x - sn * y, c[1] + sn * x + cs * y) for x, y in pts]

    def clip(subject, clipper):
        out = subject
        for i in range(len(clipper)):
            a, b = clipper[i], clipper[(i + 1) % len(clipper)]
            inp, out = out, []
            if not inp:
                break
            side = lambda p: (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])
            for k in range(len(inp)):
                p, q = inp[k], inp[(k + 1) % len(inp)]
                sp, sq = side(p), side(q)
                if sp >= 0:
                    out.append(p)
                if (sp >= 0) != (sq >= 0):
                    t = sp / (sp - sq)
                    out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
        return out

    def area(poly):
        return 0.5 * abs(sum(poly[i][0] * poly[(i + 1) % len(poly)][1] - poly[(i + 1) % len(poly)][0] * poly[i][1]
                             for i in range(len(poly)))) if len(poly) >= 3 else 0.0

    rng = np.random.default_rng(5)
    n_hit = 0
    for _ in range(4000):
        a = (rng.uniform(-1, 1), rng.uniform(-1, 1)); b = (rng.uniform(-1, 1), rng.uniform(-1, 1))
        ea, eb = rng.uniform(0.1, 1.2, 2), rng.uniform(0.1, 1.2, 2)
        ya, yb = rng.uniform(-3.2, 3.2), rng.uniform(-3.2, 3.2)
        ar = area(clip(corners(a, ea, ya), corners(b, eb, yb)))
        if 1e-9 < ar or ar == 0.0:                                  # skip numerically marginal overlaps
            hit = chk.boxes_collide(a, ea, ya, b, eb, yb)
            if ar > 1e-9:
                assert hit
                n_hit += 1
            elif hit:                                               # SAT says touching/overlap, clipping found nothing: must be marginal
                ar2 = area(clip(corners(a, ea * (1 + 1e-6), ya), corners(b, eb * (1 + 1e-6), yb)))
                assert ar2 >= 0.0
    assert n_hit > 500


def test_check_scene_semantics():
    pool = scenes.qualitative_batch(16, 4)
    gt = pool.x[:, 2:6].numpy().copy()
    solved, ncol, nmiss = chk.check_batch(gt, pool, (2, 6))
    assert solved.all()
    # NaN rows are skipped = unsolved (ddpm.py:644-645)
    bad = gt.copy(); bad[1, 0] = np.nan
    s2, _, _ = chk.check_batch(bad, pool, (2, 6))
    assert not s2[0] and s2[1:].all()
    # predicted poses are clamped to [-1, 1] before the check (ddpm.py:620): x = 5 becomes 1 -> the tile sticks out of the tray
    bad = gt.copy(); bad[2, 0] = 5.0
    s3, c3, _ = chk.check_batch(bad, pool, (2, 6))
    assert not s3[0] and c3[0] > 0
    # moving every tile of scene 1 onto the same spot collides
    bad = gt.copy(); bad[6:10, :2] = 0.0
    s4, c4, _ = chk.check_batch(bad, pool, (2, 6))
    assert not s4[1] and c4[1] > 0 and s4[0]
    # boxes world (4-feature rows): collisions only
    b = scenes.make_batch('boxes', 8, 6, seed=1)
    s5, c5, m5 = chk.check_batch(b.x[:, 2:4].numpy(), b, (2, 4), qualitative=False)
    assert s5.all() and (m5 == 0).all()
    z = b.x[:, 2:4].numpy().copy(); z[1:7] = 0.0
    s6, _, _ = chk.check_batch(z, b, (2, 4), qualitative=False)
    assert not s6[0] and s6[1:].all()
