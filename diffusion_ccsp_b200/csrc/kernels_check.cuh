// kernels_check.cuh — k_check_solved: the per-graph success check of Trainer.evaluate for the 2-D box
// worlds, one warp per scene, whole batch in one launch.
//
// Reference path replaced (all of it per-graph Python on the CPU, trimesh + python-fcl):
//   networks/ddpm.py:620-713      clamp to [-1,1], get_all_features, per-graph loop, NaN skip, success = no evaluations
//   envs/data_utils.py:221-299    render_world_from_graph.get_node (de-normalise rows; `w,l,x,y,sn,cs` unpack at :255)
//   envs/data_utils.py:360-364    yaw_from_sn_cs
//   envs/worlds.py:662-712        construct_scene_from_graph_data;  envs/mesh_utils.py:174-191 create_tray (t = 0.1, h = 0.01)
//   envs/collisions.py:58-130     all-pairs box-box narrow phase (FCL boxBox2 == SAT, touching collides), +yaw, unrotated extents
//   envs/worlds.py:380-388, 398   pairs with `bottom` and the four wall corner pairs are ignored
//   envs/data_utils.py:427-621    compute_qualitative_constraints (13 relation types re-derived from the layout)
//   envs/worlds.py:734-764        check_constraints_satisfied: every given constraint must be present (symmetric relations in
//                                 either order, data_utils.py:418-424)
//
// Arithmetic: IEEE double with the reference's operation order and NO fused multiply-add (explicit _rn intrinsics wherever a
// product feeds a sum), so every threshold comparison sees the same doubles as the Python code; the only functions that are not
// correctly rounded on both sides are atan2 / sin / cos (<= 2 ulp), which can flip a decision only on a knife edge.
// The work is tiny (<= 32 objects per scene) and latency-bound: no tensor cores, no staging — coalescing does not matter at
// 55 KB of input per 1024 scenes; what matters is that the whole batch is ONE launch with no host round trip.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ccsp {

enum { CHECK_WORLD_BOXES = 0, CHECK_WORLD_QUALITATIVE = 1 };
constexpr int CHECK_MAX_OBJ = 32;          // 4 walls + up to 28 tiles (the reference renders at most 12, mesh_utils.py:47)
constexpr int CHECK_WARPS = 4;

struct CheckArgs {
  int kind, S, F, P, pose_begin, clamp;
  const float *x, *poses, *world_dims;
  const int *scene_node_ptr, *scene_edge_ptr, *edge_a, *edge_b, *edge_type;
  unsigned char *solved;
  int *counts;                             // [S,2]: collisions, missing constraints (-1,-1: NaN rows / malformed scene)
};

namespace chk {
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }

// `in_x_range` / `in_y_range` (data_utils.py:509-521, 543-554)
__device__ __forceinline__ bool axis_range(double lo1, double hi1, double w1, double lo2, double hi2, double w2, double thr) {
  if ((lo2 <= lo1 && lo1 < hi1 && hi1 <= hi2) || (lo1 <= lo2 && lo2 < hi2 && hi2 <= hi1)) return true;
  double overlap = 0.0;
  if (lo2 <= lo1 && lo1 <= hi2 && hi2 <= hi1) overlap = sub(hi2, lo1);
  else if (lo1 <= lo2 && lo2 <= hi1 && hi1 <= hi2) overlap = sub(hi1, lo2);
  return overlap > mul(fmin(w1, w2), thr);
}

struct Obj {            // one object of the scene (walls and tiles), held by the lane of the same slot
  double cx, cy;        // centre
  double ex, ey;        // full extents, unrotated (what FCL gets)
  double c, s;          // rotation about z handed to FCL
  double lx, ly;        // extents as the labeller sees them (swapped near +-pi/2)
};
}  // namespace chk

__global__ void __launch_bounds__(CHECK_WARPS * 32) k_check_solved(CheckArgs A) {
  using namespace chk;
  __shared__ double sh[CHECK_WARPS][8][CHECK_MAX_OBJ];
  __shared__ unsigned rel[CHECK_WARPS][6][CHECK_MAX_OBJ];      // link, leftof, topof, close, valigned, haligned: rel[.][v] bit u
  __shared__ int una[CHECK_WARPS][5][CHECK_MAX_OBJ];           // center-in, left-in, right-in, bottom-in, top-in counts per index
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sc = blockIdx.x * CHECK_WARPS + warp;
  if (sc >= A.S) return;
  const unsigned FULL = 0xffffffffu;
  const int n0 = A.scene_node_ptr[sc], n1 = A.scene_node_ptr[sc + 1];
  const int N = n1 - n0 - 1;                 // tiles (node 0 of the scene is the container)
  const int M = N + 4;                       // + 4 walls
  const bool qual = A.kind == CHECK_WORLD_QUALITATIVE;
  if (N < 0 || M > CHECK_MAX_OBJ) {
    if (lane == 0) { A.solved[sc] = 0; if (A.counts) { A.counts[2 * sc] = -1; A.counts[2 * sc + 1] = -1; } }
    return;
  }
  double (*o)[CHECK_MAX_OBJ] = sh[warp];
  for (int k = 0; k < 6; ++k) rel[warp][k][lane] = 0u;
  for (int k = 0; k < 5; ++k) una[warp][k][lane] = 0;

  // ---- full feature row of a node: [x[:, :pose_begin], clamp(poses), x[:, pose_begin+P:]]   (ddpm.py:620, 807-821)
  auto feat = [&](int node, int col) -> float {
    if (col >= A.pose_begin && col < A.pose_begin + A.P) {
      float v = A.poses[(size_t)node * A.P + (col - A.pose_begin)];
      if (A.clamp) v = v < -1.f ? -1.f : (v > 1.f ? 1.f : v);          // NaN stays NaN, like torch.clamp_
      return v;
    }
    return A.x[(size_t)node * A.F + col];
  };
  // NaN anywhere in the scene's rows -> the graph is skipped, i.e. not solved (ddpm.py:644-645)
  bool has_nan = false;
  for (int i = lane; i < (n1 - n0) * A.F; i += 32) has_nan |= isnan(feat(n0 + i / A.F, i % A.F));
  has_nan = __any_sync(FULL, has_nan);

  const double w_tray = (double)A.world_dims[2 * sc], l_tray = (double)A.world_dims[2 * sc + 1];
  // container row: w, l -> the world's size (data_utils.py:250-253; worlds.py:664, 194-196)
  const double W = mul((double)feat(n0, 0), w_tray), L = mul((double)feat(n0, 1), l_tray);
  const double t = 0.1, h = 0.01;
  (void)h;
  Obj me;
  me.cx = me.cy = me.ex = me.ey = 0.0; me.c = 1.0; me.s = 0.0; me.lx = me.ly = 0.0;
  bool bad = false;
  if (lane < M) {
    if (lane == 0) { me.ex = W; me.ey = t; me.cy = add(L, t) / 2.0; }                     // north
    else if (lane == 1) { me.ex = W; me.ey = t; me.cy = -(add(L, t)) / 2.0; }             // south
    else if (lane == 2) { me.ex = t; me.ey = add(L, mul(2.0, t)); me.cx = -(add(W, t)) / 2.0; }   // west
    else if (lane == 3) { me.ex = t; me.ey = add(L, mul(2.0, t)); me.cx = add(W, t) / 2.0; }      // east
    else {
      const int node = n0 + 1 + (lane - 4);
      me.ex = mul((double)feat(node, 0), w_tray);
      me.ey = mul((double)feat(node, 1), l_tray);
      me.cx = mul((double)feat(node, 2), w_tray) / 2.0;
      me.cy = mul((double)feat(node, 3), l_tray) / 2.0;
    }
    me.lx = me.ex; me.ly = me.ey;
    if (lane >= 4 && qual) {
      const int node = n0 + 1 + (lane - 4);
      double sn = (double)feat(node, 4), cs = (double)feat(node, 5);                      // sic (data_utils.py:255)
      const double total = sqrt(add(mul(sn, sn), mul(cs, cs)));
      sn = sn / total; cs = cs / total;
      const double yaw = atan2(sn, cs);
      bad = isnan(yaw);
      if (fabs(sub(fabs(yaw), 3.141592653589793 / 2.0)) < 0.1) { me.lx = me.ey; me.ly = me.ex; }   // :457-460
      const double qw = cos(yaw / 2.0), qz = sin(yaw / 2.0), tz = mul(2.0, qz);
      me.c = sub(1.0, mul(tz, qz)); me.s = mul(tz, qw);
    }
    o[0][lane] = me.cx; o[1][lane] = me.cy; o[2][lane] = me.ex; o[3][lane] = me.ey;
    o[4][lane] = me.c; o[5][lane] = me.s; o[6][lane] = me.lx; o[7][lane] = me.ly;
  }
  bad = __any_sync(FULL, bad);
  __syncwarp();

  const double scale = fmin(W / 3.0, L / 2.0);                                             // worlds.py:226
  const double alignment = mul(0.05, scale), farness = mul(0.5, scale), closeness = mul(0.3, scale),
               touching = mul(0.1, scale), overlap_thr = mul(0.6, scale);
  const int my_idx = lane < 4 ? 0 : lane - 3;
  const double l2 = sub(me.cx, me.lx / 2.0), r2 = add(me.cx, me.lx / 2.0);
  const double b2 = sub(me.cy, me.ly / 2.0), t2 = add(me.cy, me.ly / 2.0);
  unsigned (*R)[CHECK_MAX_OBJ] = rel[warp];

  if (lane < M && qual) {      // unary relations of this object (data_utils.py:469-478); walls count towards index 0
    if (sqrt(add(mul(me.cx, me.cx), mul(me.cy, me.cy))) < closeness) atomicAdd(&una[warp][0][my_idx], 1);
    if (r2 < 0.0) atomicAdd(&una[warp][1][my_idx], 1);
    if (l2 > 0.0) atomicAdd(&una[warp][2][my_idx], 1);
    if (t2 < 0.0) atomicAdd(&una[warp][3][my_idx], 1);
    if (b2 > 0.0) atomicAdd(&una[warp][4][my_idx], 1);
  }

  // ---- all pairs (a, lane) with a < lane in the reference's object order: north, south, west, east, tile_0 ...
  int ncol = 0;
  for (int a = 0; a < M; ++a) {
    if (lane <= a || lane >= M) continue;
    const double ax = o[0][a], ay = o[1][a], aex = o[2][a], aey = o[3][a], ac = o[4][a], as = o[5][a];
    // collision: skipped for the wall corner pairs (worlds.py:398); north-south and west-east are tested like any pair
    const bool corner = lane < 4 && ((a == 0 && lane >= 2) || (a == 1 && lane >= 2));
    if (!corner) {
      // 2-D SAT, box 1 = a, box 2 = this lane  (FCL boxBox2 restricted to the plane; an axis separates iff s > 0)
      const double px = sub(me.cx, ax), py = sub(me.cy, ay);
      const double A0 = aex / 2.0, A1 = aey / 2.0, B0 = me.ex / 2.0, B1 = me.ey / 2.0;
      const double ca = ac, sa = as, cb = me.c, sb = me.s;
      const double q00 = fabs(add(mul(ca, cb), mul(sa, sb))), q01 = fabs(add(mul(-ca, sb), mul(sa, cb)));
      const double q10 = fabs(add(mul(-sa, cb), mul(ca, sb))), q11 = fabs(add(mul(sa, sb), mul(ca, cb)));
      const double pp0 = add(mul(ca, px), mul(sa, py)), pp1 = add(mul(-sa, px), mul(ca, py));
      const double t0 = add(mul(cb, px), mul(sb, py)), t1 = add(mul(-sb, px), mul(cb, py));
      bool sep = sub(fabs(pp0), add(add(A0, mul(B0, q00)), mul(B1, q01))) > 0.0;
      sep = sep || sub(fabs(pp1), add(add(A1, mul(B0, q10)), mul(B1, q11))) > 0.0;
      sep = sep || sub(fabs(t0), add(add(mul(A0, q00), mul(A1, q10)), B0)) > 0.0;
      sep = sep || sub(fabs(t1), add(add(mul(A0, q01), mul(A1, q11)), B1)) > 0.0;
      ncol += !sep;
    }
    if (!qual) continue;
    const int p = a < 4 ? 0 : a - 3, q = my_idx;
    if (p == q) continue;                                            // two walls (data_utils.py:486-487)
    const double alx = o[6][a], aly = o[7][a];
    const double l1 = sub(ax, alx / 2.0), r1 = add(ax, alx / 2.0), b1 = sub(ay, aly / 2.0), t1 = add(ay, aly / 2.0);
    if (p != 0 && q != 0) {                                          // :499-503
      if (fabs(sub(ax, me.cx)) < alignment) { atomicOr(&R[4][p], 1u << q); atomicOr(&R[4][q], 1u << p); }
      if (fabs(sub(ay, me.cy)) < alignment) { atomicOr(&R[5][p], 1u << q); atomicOr(&R[5][q], 1u << p); }
    }
    // gap entries: (relation table, upper/left index u, lower/right index v, d)                   :523-573
    auto gap = [&](int table, int u, int v, double d) {
      if (!(-0.05 <= d && d < farness)) return;
      atomicOr(&R[0][p], 1u << q); atomicOr(&R[0][q], 1u << p);      // neighbours of each other
      if (p == 0 || q == 0) return;
      if (d < closeness) atomicOr(&R[table][v], 1u << u);            // left-of(u, v) / top-of(u, v)    :586-591
      if (d < touching) { atomicOr(&R[3][p], 1u << q); atomicOr(&R[3][q], 1u << p); }   // close-to     :592-594
    };
    if (axis_range(l1, r1, alx, l2, r2, me.lx, overlap_thr)) {
      gap(2, q, p, sub(b2, t1));       // this lane's object above a
      gap(2, p, q, sub(b1, t2));       // a above this lane's object
    }
    if (axis_range(b1, t1, aly, b2, t2, me.ly, overlap_thr)) {
      gap(1, q, p, sub(l1, r2));       // this lane's object left of a
      gap(1, p, q, sub(l2, r1));       // a left of this lane's object
    }
  }
  for (int off = 16; off; off >>= 1) ncol += __shfl_xor_sync(FULL, ncol, off);
  __syncwarp();

  // ---- every given constraint must be present                                                   worlds.py:748-753
  int nmiss = 0;
  if (qual) {
    if (lane <= N) {                  // left-in/right-in and bottom-in/top-in on the same index cancel one for one (:606-613)
      int both = min(una[warp][1][lane], una[warp][2][lane]);
      una[warp][1][lane] -= both; una[warp][2][lane] -= both;
      both = min(una[warp][3][lane], una[warp][4][lane]);
      una[warp][3][lane] -= both; una[warp][4][lane] -= both;
    }
    __syncwarp();
    const int e0 = A.scene_edge_ptr[sc], e1 = A.scene_edge_ptr[sc + 1];
    for (int e = e0 + lane; e < e1; e += 32) {
      const int typ = A.edge_type[e], a = A.edge_a[e], b = A.edge_b[e];
      if (typ >= 13) continue;                                       // data_utils.py:180-181
      const bool a_tile = a >= 1 && a <= N, b_tile = b >= 1 && b <= N;
      const bool a_idx = a >= 0 && a <= N;
      bool present = false;
      switch (typ) {
        case 0: present = a_tile && b == 0; break;                                   // in        (worlds.py:137)
        case 1: present = a_idx && b == 0 && una[warp][0][a] > 0; break;             // center-in
        case 2: present = a_idx && b == 0 && una[warp][1][a] > 0; break;             // left-in
        case 3: present = a_idx && b == 0 && una[warp][2][a] > 0; break;             // right-in
        case 4: present = a_idx && b == 0 && una[warp][4][a] > 0; break;             // top-in
        case 5: present = a_idx && b == 0 && una[warp][3][a] > 0; break;             // bottom-in
        case 6: present = a_tile && b_tile && a != b; break;                         // cfree     (worlds.py:138-144, either order)
        case 7: present = a_tile && b_tile && ((R[1][b] >> a) & 1u); break;          // left-of(a, b)
        case 8: present = a_tile && b_tile && ((R[2][b] >> a) & 1u); break;          // top-of(a, b)
        case 9: present = a_tile && b_tile && ((R[3][b] >> a) & 1u); break;          // close-to
        case 10: present = a_tile && b_tile && a != b && !((R[0][b] >> a) & 1u); break;   // away-from: no gap entry at all (:598-600)
        case 11: present = a_tile && b_tile && ((R[5][b] >> a) & 1u); break;         // h-aligned
        case 12: present = a_tile && b_tile && ((R[4][b] >> a) & 1u); break;         // v-aligned
        default: present = false; break;                                            // negative ids: never present
      }
      // len(missing) of worlds.py:753: both lists are expanded first, so a symmetric relation counts twice (data_utils.py:418-424)
      if (!present) nmiss += (typ == 6 || (typ >= 9 && typ <= 12)) ? 2 : 1;
    }
    for (int off = 16; off; off >>= 1) nmiss += __shfl_xor_sync(FULL, nmiss, off);
  }
  if (lane == 0) {
    if (has_nan || bad) {
      A.solved[sc] = 0;
      if (A.counts) { A.counts[2 * sc] = -1; A.counts[2 * sc + 1] = -1; }
    } else {
      if (ncol > 0) nmiss = 0;                                       // collisions are reported alone (worlds.py:738-746)
      A.solved[sc] = (ncol == 0 && nmiss == 0) ? 1 : 0;
      if (A.counts) { A.counts[2 * sc] = ncol; A.counts[2 * sc + 1] = nmiss; }
    }
  }
}

}  // namespace ccsp
