"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's per-graph success check.

Only tests/, __graft_entry__.smoke() and bench.py's checker legs may import this module; the product path
(diffusion_ccsp_b200/checker.py -> ccsp_check_solved, CUDA) never does.

What is restated, literally and in Python floats (IEEE double, like the reference):

  denormalise_rows           envs/data_utils.py:221-299   get_node (4- and 6-feature rows; note the `w,l,x,y,sn,cs`
                                                          unpack at :255 although rows hold [x,y,cs,sn]) and
                             envs/data_utils.py:360-364   yaw_from_sn_cs
  scene_objects              envs/mesh_utils.py:174-191   create_tray (bottom + 4 walls, t = 0.1), envs/worlds.py:42-46 (h = 0.01),
                             envs/worlds.py:662-712       construct_scene_from_graph_data (tile_box_{i-1}, rotations[...] = yaw),
                             envs/worlds.py:147-196       generate_json (center = centroid, extents = the UNROTATED box extents)
  box_collisions             envs/collisions.py:58-130    all-pairs fcl.collide over Box bodies, transform = (+yaw about z, centroid);
                             envs/worlds.py:380-388, 398  drop pairs with 'bottom' and the four wall-wall corner pairs
  qualitative_relations      envs/data_utils.py:427-621   compute_qualitative_constraints restated as relation tables (the SET of
                                                          relations, symmetric ones in both orders); pinned against the
                                                          reference's own function on thousands of random and near-threshold
                                                          layouts by tests/test_checker_oracle.py
  check_scene                envs/worlds.py:734-764       check_constraints_satisfied (+ generate_constraints :125-145,
                             envs/data_utils.py:173-186, 418-424 constraint_from_edge_attr / expand_unordered_constraints)
  check_batch                networks/ddpm.py:620-713     clamp, get_all_features, per-graph loop, NaN skip, `success = no evaluations`

Third-party arithmetic that is absent from /root/reference: python-fcl (requirements.txt:3, unpinned; wraps FCL >= 0.6) does the
narrow phase.  For two boxes FCL runs `boxBox2` (the ODE dBoxBox separating-axis test: 3 + 3 face axes and 9 edge-edge axes, an axis
separates iff |t . axis| - (r_A + r_B) > 0, so touching boxes collide).  All bodies here share the same z extent and z centre and are
rotated about z only, so the edge-edge axes are either degenerate or repeat a face axis with a 1e-6 fudge ADDED to the radii (they can
never separate what the face axes did not) and the z face axis always overlaps: the test reduces to the 2-D SAT on the four face
normals restated in `boxes_collide`.  FCL itself cannot be run here => the collision half is pinned against this restatement and
hand-built touching / overlapping / rotated cases, NOT against FCL ("parity unpinned" for that half; DESIGN.md says so).
trimesh (centroid of the transformed mesh) is also absent: centres are taken as the exact (x, y, h/2) handed to the transform,
i.e. without trimesh's O(1e-16) round-off.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

qualitative_constraint_names = [            # networks/denoise_fn.py:19-23
    'in', 'center-in', 'left-in', 'right-in', 'top-in', 'bottom-in',
    'cfree', 'left-of', 'top-of',
    'close-to', 'away-from', 'h-aligned', 'v-aligned'
]
ignored_constraints = ['right-of', 'bottom-of']      # networks/denoise_fn.py:25

TRAY_T = 0.1        # envs/worlds.py:389 (TrayWorld t)
WORLD_H = 0.01      # envs/worlds.py:46 (orthographic)
WALL_CFREE = [('north', 'east'), ('south', 'east'), ('north', 'west'), ('south', 'west')]   # envs/worlds.py:398


# ------------------------------------------------------------------------------------------------------
def yaw_from_sn_cs(sn, cs):
    """envs/data_utils.py:360-364 (np.sqrt / np.arctan2 on Python floats)."""
    total = np.sqrt(sn ** 2 + cs ** 2)
    with np.errstate(all='ignore'):
        sn = sn / total
        cs = cs / total
    return float(np.arctan2(sn, cs))


def denormalise_rows(features: Sequence[Sequence[float]], world_dims=(3, 2)) -> np.ndarray:
    """envs/data_utils.py:225-299 for 4- and 6-feature rows: returns the `nodes` array [type, w, l, x, y(, yaw)]."""
    w_tray, l_tray = world_dims
    nodes = []
    for i, f in enumerate(features):
        f = [float(v) for v in f]
        typ = int(i != 0)
        if len(f) == 4:
            w, l, x, y = f
            geom = [w * w_tray, l * l_tray]
            pose = [x * w_tray / 2, y * l_tray / 2]
        elif len(f) == 6:
            if w_tray == 3 and l_tray == 3:
                raise NotImplementedError('triangle P1 encoding (data_utils.py:239-247) is not a box world')
            if typ == 0:
                w, l, x, y, _, _ = f
                geom = [w * w_tray, l * l_tray]
                pose = [x * w_tray / 2, y * l_tray / 2, 0]
            else:
                w, l, x, y, sn, cs = f                     # sic: rows hold [.., cs, sn]   (data_utils.py:255)
                roll = yaw_from_sn_cs(sn, cs)
                geom = [w * w_tray, l * l_tray]
                pose = [x * w_tray / 2, y * l_tray / 2, roll]
        else:
            raise NotImplementedError(f'{len(f)}-feature rows are not a 2-D box world')
        nodes.append([typ] + geom + pose)
    return np.asarray(nodes, dtype=np.float64)


def scene_objects(nodes: np.ndarray, qualitative: bool):
    """objects (label -> center / extents, insertion order = the reference's dict order) and rotations."""
    w, l = nodes[0, 1:3]
    w, l = float(w), float(l)
    h, t = WORLD_H, TRAY_T
    objects = {
        'bottom': {'extents': (w, l, t), 'center': (0, 0, -t / 2)},
        'north': {'extents': (w, t, h), 'center': (0, (l + t) / 2, h / 2)},
        'south': {'extents': (w, t, h), 'center': (0, -(l + t) / 2, h / 2)},
        'west': {'extents': (t, l + 2 * t, h), 'center': (-(w + t) / 2, 0, h / 2)},
        'east': {'extents': (t, l + 2 * t, h), 'center': ((w + t) / 2, 0, h / 2)},
    }
    rotations = {} if qualitative else None              # worlds.py:58 (None) / :724 ({})
    for i in range(1, nodes.shape[0]):
        if nodes.shape[1] == 5:
            _, bw, bl, x, y = nodes[i]
            yaw = None
        else:
            _, bw, bl, x, y, yaw = nodes[i]
            if qualitative:
                rotations[f'tile_box_{i - 1}'] = float(yaw)
        objects[f'tile_box_{i - 1}'] = {'extents': (float(bw), float(bl), h), 'center': (float(x), float(y), h / 2)}
    return objects, rotations


# ------------------------------------------------------------------------------------------------------
def _rot2(yaw: float):
    """rotation about z the way the reference hands it to FCL: transformations.quaternion_about_axis(yaw, (0,0,1)) =
    (w, 0, 0, z) = (cos(yaw/2), 0, 0, sin(yaw/2)), turned into a matrix (Eigen toRotationMatrix): c = 1 - 2 z z, s = 2 w z."""
    qw, qz = math.cos(yaw / 2.0), math.sin(yaw / 2.0)
    tz = 2.0 * qz
    return 1.0 - tz * qz, tz * qw          # (cos, sin)


def boxes_collide(c1, e1, yaw1, c2, e2, yaw2) -> bool:
    """2-D SAT of two rectangles (centre, full extents, yaw about z); touching counts as a collision (FCL boxBox2: an axis
    separates iff s > 0)."""
    ca, sa = (1.0, 0.0) if yaw1 is None else _rot2(yaw1)
    cb, sb = (1.0, 0.0) if yaw2 is None else _rot2(yaw2)
    px, py = c2[0] - c1[0], c2[1] - c1[1]
    A0, A1 = e1[0] / 2.0, e1[1] / 2.0
    B0, B1 = e2[0] / 2.0, e2[1] / 2.0
    # R = R1^T R2 restricted to the plane; Q = |R|
    r00 = ca * cb + sa * sb
    r01 = -ca * sb + sa * cb
    r10 = -sa * cb + ca * sb
    r11 = sa * sb + ca * cb
    q00, q01, q10, q11 = abs(r00), abs(r01), abs(r10), abs(r11)
    # axes of box 1: pp = R1^T p
    pp0 = ca * px + sa * py
    pp1 = -sa * px + ca * py
    if abs(pp0) - (A0 + B0 * q00 + B1 * q01) > 0:
        return False
    if abs(pp1) - (A1 + B0 * q10 + B1 * q11) > 0:
        return False
    # axes of box 2: p . R2[:, i]
    t0 = cb * px + sb * py
    t1 = -sb * px + cb * py
    if abs(t0) - (A0 * q00 + A1 * q10 + B0) > 0:
        return False
    if abs(t1) - (A0 * q01 + A1 * q11 + B1) > 0:
        return False
    return True


def box_collisions(objects: Dict[str, dict], rotations: Optional[Dict[str, float]]) -> List[Tuple[str, str]]:
    """envs/collisions.py:58-130 + the filters of envs/worlds.py:380-388.  All bodies of these worlds are boxes whose z ranges
    overlap pairwise EXCEPT bottom (z in [-t, 0] vs [0, h]: touching => FCL collides), and every pair with bottom is dropped
    by the filter anyway, so bottom is skipped."""
    labels = [k for k in objects if k != 'bottom']
    out = []
    for a in range(len(labels)):
        for b in range(a + 1, len(labels)):          # (i, j) is recorded once, first-seen order (collisions.py:124-125)
            i, j = labels[a], labels[b]
            if (i, j) in WALL_CFREE:
                continue
            yi = rotations.get(i) if rotations is not None else None
            yj = rotations.get(j) if rotations is not None else None
            if boxes_collide(objects[i]['center'], objects[i]['extents'], yi, objects[j]['center'], objects[j]['extents'], yj):
                out.append((i, j))
    return out


# ------------------------------------------------------------------------------------------------------
def _bounds(obj, name, rotations):
    """(left, right, bottom, top, cx, cy, lx, ly) of an object's box as the labeller sees it: extents are swapped iff the
    object's yaw is within 0.1 of +-pi/2 (data_utils.py:457-460, 491-494)."""
    x, y, _ = obj['center']
    lx, ly, _ = obj['extents']
    if rotations is not None and name in rotations:
        if abs(abs(rotations[name]) - np.pi / 2) < 0.1:
            ly, lx, _ = obj['extents']
    return x - lx / 2, x + lx / 2, y - ly / 2, y + ly / 2, x, y, lx, ly


def _axis_range(lo1, hi1, w1, lo2, hi2, w2, overlap_threshold):
    """`in_x_range` / `in_y_range` of data_utils.py:509-521, 543-554: containment, or a partial overlap larger than
    overlap_threshold x the smaller width."""
    if (lo2 <= lo1 < hi1 <= hi2) or (lo1 <= lo2 < hi2 <= hi1):
        return True
    overlap = 0
    if lo2 <= lo1 <= hi2 <= hi1:
        overlap = hi2 - lo1
    elif lo1 <= lo2 <= hi1 <= hi2:
        overlap = hi1 - lo2
    return overlap > min(w1, w2) * overlap_threshold


def qualitative_relations(objects, rotations=None, scale=1):
    """The SET of relations envs/data_utils.py:427-621 (compute_qualitative_constraints) derives from a layout, restated as
    relation tables instead of the reference's lists (this is also the formulation of the CUDA kernel):

      index 0 = 'bottom' and every tray wall (walls are aliases of bottom, :453, :482), index i = i-th tile;
      unary    cnt[rel][a]     how many objects of index a satisfy center-in / left-in / right-in / bottom-in / top-in
                               (:469-478); a left-in/right-in (bottom-in/top-in) pair on the same index cancels one-for-one
                               (:606-613) — only index 0 can ever hold both;
      gap      pairs (p above q, d) and (p left of q, d) for every unordered pair of distinct indices with enough overlap
               on the other axis and -0.05 <= d < farness (:523-573);
      then     top-of(p, q) / left-of(p, q)   iff such a gap with d < closeness, p, q != 0           (:586-591)
               close-to{p, q}                 iff any gap between them with d < touching, p, q != 0  (:592-594)
               away-from{p, q}                iff p, q != 0 tiles with NO gap entry at all           (:598-600)
               v-/h-aligned{p, q}             iff |dx| / |dy| < alignment, both tiles               (:499-503)
    Symmetric relations are emitted in both orders (callers compare after expand_unordered_constraints, :418-424)."""
    alignment, farness, closeness, touching, overlap_threshold = (0.05 * scale, 0.5 * scale, 0.3 * scale, 0.1 * scale,
                                                                  0.6 * scale)
    names = [n for n in objects if n != 'bottom']
    tile_names = [n for n in objects if 'tile_' in n]
    index = {n: (1 + tile_names.index(n) if 'tile_' in n else 0) for n in names}
    n_idx = 1 + len(tile_names)
    box = {n: _bounds(objects[n], n if 'tile_' in n else 'bottom', rotations) for n in names}

    unary = {k: [0] * n_idx for k in ('center-in', 'left-in', 'right-in', 'bottom-in', 'top-in')}
    for n in names:
        left, right, bottom, top, cx, cy, _, _ = box[n]
        a = index[n]
        unary['center-in'][a] += math.sqrt(cx ** 2 + cy ** 2) < closeness
        unary['left-in'][a] += right < 0
        unary['right-in'][a] += left > 0
        unary['bottom-in'][a] += top < 0
        unary['top-in'][a] += bottom > 0
    for ka, kb in (('left-in', 'right-in'), ('bottom-in', 'top-in')):
        for a in range(n_idx):
            both = min(unary[ka][a], unary[kb][a])
            unary[ka][a] -= both
            unary[kb][a] -= both

    out = set()
    for k, cnt in unary.items():
        out |= {(k, a, 0) for a in range(n_idx) if cnt[a] > 0}

    linked = set()                       # unordered index pairs with any gap entry (the `neighbor` lists, :596)
    for i, n1 in enumerate(names):
        l1, r1, b1, t1, x1, y1, w1, h1 = box[n1]
        for n2 in names[i + 1:]:
            p, q = index[n1], index[n2]
            if p == q:                   # two walls
                continue
            l2, r2, b2, t2, x2, y2, w2, h2 = box[n2]
            if p != 0 and q != 0:
                if abs(x1 - x2) < alignment:
                    out |= {('v-aligned', p, q), ('v-aligned', q, p)}
                if abs(y1 - y2) < alignment:
                    out |= {('h-aligned', p, q), ('h-aligned', q, p)}
            gaps = []                    # (relation, upper/left index, lower/right index, d)
            if _axis_range(l1, r1, w1, l2, r2, w2, overlap_threshold):
                gaps.append(('top-of', q, p, b2 - t1))        # n2 above n1
                gaps.append(('top-of', p, q, b1 - t2))        # n1 above n2
            if _axis_range(b1, t1, h1, b2, t2, h2, overlap_threshold):
                gaps.append(('left-of', q, p, l1 - r2))       # n2 left of n1
                gaps.append(('left-of', p, q, l2 - r1))       # n1 left of n2
            for rel, u, v, d in gaps:
                if not (-0.05 <= d < farness):
                    continue
                linked.add((min(p, q), max(p, q)))
                if p == 0 or q == 0:
                    continue
                if d < closeness:
                    out.add((rel, u, v))
                if d < touching:
                    out |= {('close-to', p, q), ('close-to', q, p)}
    for p in range(1, n_idx):
        for q in range(p + 1, n_idx):
            if (p, q) not in linked:
                out |= {('away-from', p, q), ('away-from', q, p)}
    return out


def qualitative_constraints(objects, rotations=None, scale=1):
    return sorted(qualitative_relations(objects, rotations, scale))


def expand_unordered_constraints(constraints):
    """envs/data_utils.py:418-424."""
    new_constraints = []
    for c in constraints:
        if c[0] in ['close-to', 'away-from', 'h-aligned', 'v-aligned', 'cfree']:
            new_constraints.append(tuple([c[0], c[2], c[1]]))
        new_constraints.append(c)
    return new_constraints


def constraint_from_edge_attr(edge_attr, edge_index) -> List[list]:
    """envs/data_utils.py:173-186 (composed_inference=False): [name, a, b]; type ids beyond the vocabulary are skipped."""
    out = []
    for i in range(len(edge_attr)):
        typ = int(edge_attr[i])
        if typ >= len(qualitative_constraint_names):
            continue
        out.append([qualitative_constraint_names[typ]] + [int(edge_index[0][i]), int(edge_index[1][i])])
    return out


# ------------------------------------------------------------------------------------------------------
def check_scene(features, constraints=None, world_dims=(3, 2), qualitative=True, labeller=None):
    """One graph: returns (collisions, missing) — the scene is solved iff both are empty (ddpm.py:704-713).
    `labeller` replaces the port of compute_qualitative_constraints (tests pass the reference's own function)."""
    nodes = denormalise_rows(features, world_dims)
    objects, rotations = scene_objects(nodes, qualitative)
    collisions = box_collisions(objects, rotations)
    if not qualitative or len(collisions) > 0:                       # worlds.py:377-388 / :738-746
        return collisions, []
    n_obj = nodes.shape[0]                                           # bottom + tiles  (generate_constraints, worlds.py:125-145)
    current = [('in', i, 0) for i in range(1, n_obj)]
    current += [('cfree', i, j) for i in range(1, n_obj - 1) for j in range(i + 1, n_obj)]
    w, l = float(nodes[0, 1]), float(nodes[0, 2])
    scale = min([w / 3, l / 2])                                      # worlds.py:226
    current += (labeller or qualitative_constraints)(objects, rotations=rotations, scale=scale)
    current = [tuple(d) for d in current if d[0] not in ignored_constraints]
    given = [tuple(d) for d in constraints if d[0] not in ignored_constraints]
    current = expand_unordered_constraints(current)
    given = expand_unordered_constraints(given)
    missing = [ct for ct in given if ct not in current]
    return collisions, missing


def check_batch(poses: np.ndarray, batch, pose_slice: Tuple[int, int], qualitative=True, world_dims=(3, 2), labeller=None):
    """networks/ddpm.py:620-713 restricted to the success decision: clamp to [-1,1], re-assemble rows, per-graph check.
    Returns (solved bool [S], n_collisions i32 [S], n_missing i32 [S])."""
    x = np.asarray(batch.x, dtype=np.float32)
    poses = np.clip(np.asarray(poses, dtype=np.float32), -1.0, 1.0)                  # ddpm.py:620
    rows = np.concatenate([x[:, :pose_slice[0]], poses, x[:, pose_slice[1]:]], 1)     # ddpm.py:807-821
    sid = np.asarray(batch.x_extract).astype(np.int64)
    esid = np.asarray(batch.edge_extract).astype(np.int64)
    ei = np.asarray(batch.edge_index)
    ea = np.asarray(batch.edge_attr)
    S = int(sid.max()) + 1
    solved = np.zeros(S, bool)
    ncol = np.zeros(S, np.int32)
    nmiss = np.zeros(S, np.int32)
    for j in range(S):
        feats = rows[sid == j]
        if np.isnan(feats).any():                                                    # ddpm.py:644-645
            ncol[j] = nmiss[j] = -1
            continue
        cons = None
        if qualitative:
            sel = np.where(esid == j)[0]
            e = ei[:, sel]
            e = e - e.min() if e.size else e                                         # ddpm.py:690-691
            cons = constraint_from_edge_attr(ea[sel], e)
        wd = world_dims[j] if isinstance(world_dims, list) else world_dims
        col, miss = check_scene(feats, cons, wd, qualitative, labeller)
        ncol[j], nmiss[j] = len(col), len(miss)
        solved[j] = len(col) == 0 and len(miss) == 0
    return solved, ncol, nmiss
