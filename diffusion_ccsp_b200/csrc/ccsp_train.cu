// ccsp_train.cu — host side of the training step: compiled training graph, loss + gradients, Adam.
// extern "C" entry points are declared in include/ccsp_b200.h.
#include "../../include/ccsp_b200.h"

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels_train.cuh"

using namespace ccsp;
using namespace ccsp::train;

#define TR_REQUIRE(cond, msg)                                \
  do {                                                       \
    if (!(cond)) {                                           \
      set_error(std::string("invalid argument: ") + msg);    \
      return CCSP_ERR_INVALID;                               \
    }                                                        \
  } while (0)

namespace {
// Device blocks of destroyed training graphs, reused by the next one: a training loop compiles a new batch every step and
// cudaMalloc / cudaFree of ~40 blocks per step (cudaFree synchronises the device) would dominate a few-millisecond step.
struct TrainBlockCache {
  struct Blk { void *p; size_t bytes; int device; };
  std::vector<Blk> blocks;
  size_t bytes = 0;
  static constexpr size_t kCap = (size_t)6 << 30;
  void *take(size_t need, int device, size_t *actual) {
    int best = -1;
    for (int i = 0; i < (int)blocks.size(); ++i)
      if (blocks[i].device == device && blocks[i].bytes >= need && blocks[i].bytes <= 2 * need + ((size_t)1 << 16) &&
          (best < 0 || blocks[i].bytes < blocks[best].bytes))
        best = i;
    if (best < 0) return nullptr;
    void *p = blocks[best].p;
    *actual = blocks[best].bytes;
    bytes -= blocks[best].bytes;
    blocks.erase(blocks.begin() + best);
    return p;
  }
  void give(void *p, size_t sz, int device) {
    if (bytes + sz > kCap) { cudaFree(p); return; }
    blocks.push_back(Blk{p, sz, device});
    bytes += sz;
  }
};
thread_local TrainBlockCache g_train_cache;     // one per host thread / rank, like the handles themselves
}  // namespace

struct CcspTrainGraph {
  int device = 0;
  int G = 0, P = 0, Gr = 0, C = 0, F = 0, normalize = 1, pose_begin = 0, grasp_begin = 0;
  int64_t n = 0, E = 0, Epad = 0;
  int nseg = 4, Kseg = 1024, Kin = 1280;
  int start[MAX_TYPES + 1] = {0};        // padded row range per type
  int rows_of_type[MAX_TYPES] = {0};     // real edges per type
  std::vector<std::pair<void *, size_t>> blocks;
  // graph
  float *x = nullptr, *x0 = nullptr, *xtail = nullptr, *freqs = nullptr;
  int *src_i = nullptr, *src_j = nullptr, *tile_type = nullptr, *node_ptr = nullptr, *node_src = nullptr, *type_rows = nullptr;
  signed char *mask = nullptr;
  // activations (forward) and their gradients
  float *noise = nullptr, *xt = nullptr;
  float *enc_z1[3] = {}, *enc_a1[3] = {}, *enc_z2[3] = {}, *enc_e[3] = {}, *enc_dz2[3] = {}, *enc_dz1[3] = {};   // geom, pose, grasp
  float *emb = nullptr, *tz1 = nullptr, *ta1 = nullptr, *temb = nullptr, *dtemb = nullptr, *bias = nullptr, *dbias = nullptr;
  float *Z = nullptr, *H = nullptr, *D1 = nullptr, *A1 = nullptr, *O = nullptr, *out = nullptr, *err = nullptr;
  float *dOut = nullptr, *dO = nullptr, *dD1 = nullptr, *dZ = nullptr, *dIn = nullptr, *part = nullptr, *colpart = nullptr, *dtemb_part = nullptr;
  size_t part_floats = 0;

  template <typename T>
  cudaError_t alloc(T **p, size_t count) {
    size_t bytes = ((count ? count : 1) * sizeof(T) + 255) & ~(size_t)255, actual = 0;
    void *q = g_train_cache.take(bytes, device, &actual);
    cudaError_t e = cudaSuccess;
    if (!q) { actual = bytes; e = cudaMalloc(&q, bytes); }
    if (e == cudaSuccess) { blocks.emplace_back(q, actual); *p = (T *)q; }
    return e;
  }
  template <typename T>
  cudaError_t upload(T **p, const std::vector<T> &h) {
    cudaError_t e = alloc(p, h.size());
    if (e != cudaSuccess || h.empty()) return e;
    return cudaMemcpyAsync(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, upload_stream);
  }
  cudaStream_t upload_stream = nullptr;
  void free_all() {          // the caller has synchronised the device: the blocks may be reused at once
    for (auto &b : blocks) g_train_cache.give(b.first, b.second, device);
    blocks.clear();
  }
};

namespace {

template <class Prob>
int launch_gemm(const Prob &p, int M, int N, int Z, cudaStream_t st) {
  dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM), (unsigned)Z);
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return CCSP_OK;
  k_sgemm<Prob><<<grid, NT, 0, st>>>(p);
  CCSP_LAUNCH_CHECK();
  return CCSP_OK;
}

// dW[M, N] = sum_r dY[r, M]^T X[r, N], deterministic split over the R rows
int weight_grad(CcspTrainGraph *g, const float *dY, const float *X, int M, int N, int64_t R, float *dW, cudaStream_t st) {
  int slices = (int)std::min<int64_t>(64, std::max<int64_t>(1, R / 512));
  while ((size_t)slices * M * N > g->part_floats && slices > 1) --slices;
  LinearBwdWeight p;
  p.dY = dY; p.X = X; p.M_ = M; p.N_ = N; p.R_ = (int)R;
  p.chunk = (int)((R + slices - 1) / slices);
  p.part = slices == 1 ? dW : g->part;
  int rc = launch_gemm(p, M, N, slices, st);
  if (rc || slices == 1) return rc;
  const size_t count = (size_t)M * N;
  k_reduce_parts<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(g->part, slices, count, dW);
  CCSP_LAUNCH_CHECK();
  return CCSP_OK;
}

int col_sums(ColSumArgs &a, cudaStream_t st) {
  k_colsum<<<dim3((unsigned)((a.cols + 31) / 32), (unsigned)a.groups, COLSUM_SLICES), 256, 0, st>>>(a);
  CCSP_LAUNCH_CHECK();
  k_colsum_finish<<<(unsigned)((a.groups * a.cols + 255) / 256), 256, 0, st>>>(a);
  CCSP_LAUNCH_CHECK();
  return CCSP_OK;
}
int col_sum(CcspTrainGraph *g, const float *Mx, int ld, int cols, int64_t r0, int64_t r1, float *out, cudaStream_t st) {
  ColSumArgs a;
  std::memset(&a, 0, sizeof(a));
  a.M = Mx; a.ld = ld; a.cols = cols; a.groups = 1; a.start[0] = (int)r0; a.start[1] = (int)r1; a.out = out; a.part = g->colpart;
  return col_sums(a, st);
}

struct EncPtrs { const float *w0, *b0, *w2, *b2; };
struct EncGrads { float *w0, *b0, *w2, *b2; };

}  // namespace

// forward of ConstraintDiffuser on g->xt (denoise_fn.py:466-521): encoders, time term, first layer, decoder -> g->O [2 Epad, P]
static int forward_pass(CcspTrainGraph *g, const CcspParams *w, const PtrTable &W, const PtrTable &B, int t, SegSrc &in, cudaStream_t st) {
  const int n = (int)g->n, P = g->P, C = g->C;
  const int64_t Epad = g->Epad;
  const int rows2 = (int)(2 * Epad);
  const int ntab = g->Gr > 0 ? 3 : 2;
  const EncPtrs ew[3] = {{w->geom_w0, w->geom_b0, w->geom_w2, w->geom_b2}, {w->pose_w0, w->pose_b0, w->pose_w2, w->pose_b2},
                         {w->grasp_w0, w->grasp_b0, w->grasp_w2, w->grasp_b2}};
  const float *enc_in[3] = {g->x, g->xt, g->x};
  const int enc_ld[3] = {g->F, P, g->F}, enc_off[3] = {0, 0, g->grasp_begin}, enc_k[3] = {g->G, P, g->Gr};
  int rc;
  for (int tb = 0; tb < ntab; ++tb) {
    k_enc1_fwd<<<(n * CCSP_HH + 255) / 256, 256, 0, st>>>(enc_in[tb], enc_ld[tb], enc_off[tb], enc_k[tb], n, ew[tb].w0, ew[tb].b0,
                                                          g->enc_z1[tb], g->enc_a1[tb]);
    CCSP_LAUNCH_CHECK();
    LinearFwd p;
    p.X = g->enc_a1[tb]; p.W = ew[tb].w2; p.bias = ew[tb].b2; p.Z = g->enc_z2[tb]; p.Y = g->enc_e[tb];
    p.M_ = n; p.N_ = CCSP_H; p.K_ = CCSP_HH;
    if ((rc = launch_gemm(p, n, CCSP_H, 1, st))) return rc;
  }
  k_time_gemv<0><<<4 * CCSP_H / 8, 256, CCSP_H * sizeof(float), st>>>(t, g->freqs, nullptr, w->time_w1, w->time_b1, 4 * CCSP_H, CCSP_H,
                                                                     g->emb, g->tz1, g->ta1);
  CCSP_LAUNCH_CHECK();
  k_time_gemv<1><<<CCSP_H / 8, 256, 4 * CCSP_H * sizeof(float), st>>>(t, g->freqs, g->ta1, w->time_w3, w->time_b3, CCSP_H, 4 * CCSP_H,
                                                                     nullptr, nullptr, g->temb);
  CCSP_LAUNCH_CHECK();
  k_time_bias_fwd<<<C, 512, 0, st>>>(W, B, g->Kin, g->Kseg, g->temb, g->bias);
  CCSP_LAUNCH_CHECK();
  std::memset(&in, 0, sizeof(in));
  {
    int s = 0;
    if (g->Gr > 0) { in.base[s] = g->enc_e[2]; in.idx[s] = g->src_i; ++s; }      // grasp_emb[args_1]  (denoise_fn.py:337)
    in.base[s] = g->enc_e[0]; in.idx[s] = g->src_i; ++s;
    in.base[s] = g->enc_e[0]; in.idx[s] = g->src_j; ++s;
    in.base[s] = g->enc_e[1]; in.idx[s] = g->src_i; ++s;
    in.base[s] = g->enc_e[1]; in.idx[s] = g->src_j; ++s;
  }
  if (Epad > 0) {
    EdgeL1Fwd p;
    p.in = in; p.W = W; p.tile_type = g->tile_type; p.bias = g->bias; p.Z = g->Z; p.H = g->H;
    p.Epad = (int)Epad; p.Kseg = g->Kseg; p.Kin = g->Kin;
    if ((rc = launch_gemm(p, (int)Epad, CCSP_H2, 1, st))) return rc;
    LinearFwd q;
    q.X = g->H; q.W = w->dec_w0; q.bias = w->dec_b0; q.Z = g->D1; q.Y = g->A1;
    q.M_ = rows2; q.N_ = CCSP_HH; q.K_ = CCSP_H;
    if ((rc = launch_gemm(q, rows2, CCSP_HH, 1, st))) return rc;
    k_dec2_fwd<<<(rows2 + 255) / 256, 256, 0, st>>>(g->A1, w->dec_w2, w->dec_b2, rows2, P, g->O);
    CCSP_LAUNCH_CHECK();
  }
  return CCSP_OK;
}

extern "C" {

int ccsp_train_graph_create(const CcspTrainDims *d, const float *x, int64_t n, const int64_t *edge_index, const float *edge_attr,
                            const int8_t *mask, int64_t E, void *stream, CcspTrainGraph **out) {
  TR_REQUIRE(d && x && mask && out, "null argument");
  TR_REQUIRE(d->hidden_dim == CCSP_HIDDEN_DIM, "hidden_dim must be 256");
  TR_REQUIRE(d->pose_dim >= 1 && d->pose_dim <= CCSP_MAX_POSE_DIM && d->geom_dim >= 1 && d->geom_dim <= CCSP_MAXP, "pose/geom width out of range");
  TR_REQUIRE(d->grasp_dim >= 0 && d->grasp_dim <= CCSP_MAXP, "grasp width out of range");
  TR_REQUIRE(d->num_types >= 1 && d->num_types <= MAX_TYPES, "num_types out of range");
  TR_REQUIRE(n > 0 && n < (1ll << 23) && E >= 0 && E < (1ll << 22), "n/E out of range");
  TR_REQUIRE(E == 0 || (edge_index && edge_attr), "null edge arrays");
  TR_REQUIRE(d->row_width >= d->geom_dim && d->pose_begin >= 0 && d->pose_begin + d->pose_dim <= d->row_width, "feature slices exceed row width");
  TR_REQUIRE(d->grasp_dim == 0 || (d->grasp_begin >= 0 && d->grasp_begin + d->grasp_dim <= d->row_width), "grasp slice exceeds row width");
  int ndev = 0;
  CCSP_CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (ndev == 0) { set_error("no CUDA device: libccsp_b200 has no CPU fallback"); return CCSP_ERR_CUDA; }
  cudaStream_t st = (cudaStream_t)stream;
  const int C = d->num_types, P = d->pose_dim, F = d->row_width;

  // edges grouped by type (stable), every type padded to whole 64-row tiles; padded rows read the zero row n
  std::vector<int> etype(E);
  std::vector<int64_t> cnt(C, 0);
  for (int64_t e = 0; e < E; ++e) {
    const float a = edge_attr[e];                       // `edge_attr == i` on floats (denoise_fn.py:317)
    int c = (a >= 0.f && a < (float)C) ? (int)a : -1;
    if (c >= 0 && (float)c != a) c = -1;
    etype[e] = c;
    if (c >= 0) {
      const int64_t i = edge_index[e], j = edge_index[E + e];
      TR_REQUIRE(i >= 0 && i < n && j >= 0 && j < n, "edge_index out of range");
      ++cnt[c];
    }
  }
  CcspTrainGraph *g = new CcspTrainGraph();
  if (cudaGetDevice(&g->device) != cudaSuccess) { delete g; set_error("cudaGetDevice failed"); return CCSP_ERR_CUDA; }
  g->upload_stream = st;
  std::vector<int> tile_type;
  for (int c = 0; c < C; ++c) {
    const int64_t tiles = (cnt[c] + TILE_ROWS - 1) / TILE_ROWS;
    g->start[c + 1] = g->start[c] + (int)(tiles * TILE_ROWS);
    g->rows_of_type[c] = (int)cnt[c];
    for (int64_t k = 0; k < tiles; ++k) tile_type.push_back(c);
  }
  for (int c = C; c < MAX_TYPES; ++c) g->start[c + 1] = g->start[C];
  const int64_t Epad = g->start[C];
  std::vector<int> src_i(Epad, (int)n), src_j(Epad, (int)n);
  {
    std::vector<int64_t> fill(C);
    for (int c = 0; c < C; ++c) fill[c] = g->start[c];
    for (int64_t e = 0; e < E; ++e) {
      const int c = etype[e];
      if (c < 0) continue;
      const int64_t pos = fill[c]++;
      src_i[pos] = (int)edge_index[e];
      src_j[pos] = (int)edge_index[E + e];
    }
  }
  // destination CSR in the reference's accumulation order (denoise_fn.py:380-383, types visited in order at :512)
  std::vector<int> node_ptr(n + 1, 0);
  for (int64_t pos = 0; pos < Epad; ++pos)
    if (src_i[pos] < n) { ++node_ptr[src_i[pos] + 1]; ++node_ptr[src_j[pos] + 1]; }
  for (int64_t v = 0; v < n; ++v) node_ptr[v + 1] += node_ptr[v];
  std::vector<int> node_src(node_ptr[n]);
  {
    std::vector<int> fill(node_ptr.begin(), node_ptr.end() - 1);
    for (int64_t pos = 0; pos < Epad; ++pos) {
      if (src_i[pos] >= n) continue;
      node_src[fill[src_i[pos]]++] = (int)(2 * pos);
      node_src[fill[src_j[pos]]++] = (int)(2 * pos + 1);
    }
  }
  std::vector<float> x0((size_t)n * P), xtail((size_t)n * P);
  for (int64_t v = 0; v < n; ++v)
    for (int p = 0; p < P; ++p) {
      x0[v * P + p] = x[v * F + d->pose_begin + p];        // ddpm.py:367
      xtail[v * P + p] = x[v * F + (F - P) + p];           // denoise_fn.py:533
    }
  std::vector<signed char> hmask(mask, mask + n);
  for (auto &b : hmask) b = b != 0;
  std::vector<float> freqs(CCSP_HH);
  {
    const float e = (float)(-(std::log(10000.0) / (CCSP_HH - 1)));     // denoise_fn.py:45-47, FP32 like torch
    for (int k = 0; k < CCSP_HH; ++k) freqs[k] = expf((float)k * e);
  }
  std::vector<int> type_rows(g->rows_of_type, g->rows_of_type + MAX_TYPES);

  g->G = d->geom_dim; g->P = P; g->Gr = d->grasp_dim; g->C = C; g->F = F; g->normalize = d->normalize ? 1 : 0;
  g->pose_begin = d->pose_begin; g->grasp_begin = d->grasp_begin;
  g->n = n; g->E = E; g->Epad = Epad;
  g->nseg = g->Gr > 0 ? 5 : 4; g->Kseg = g->nseg * CCSP_H; g->Kin = g->Kseg + CCSP_H;
  auto fail = [&](cudaError_t e, const char *what) {
    set_error(std::string(what) + " failed: " + cudaGetErrorString(e));
    g->free_all();
    delete g;
    return CCSP_ERR_CUDA;
  };
#define G_TRY(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) return fail(_e, #expr); } while (0)
  {
    std::vector<float> hx(x, x + (size_t)n * F);
    G_TRY(g->upload(&g->x, hx));
  }
  G_TRY(g->upload(&g->x0, x0)); G_TRY(g->upload(&g->xtail, xtail)); G_TRY(g->upload(&g->freqs, freqs));
  G_TRY(g->upload(&g->src_i, src_i)); G_TRY(g->upload(&g->src_j, src_j)); G_TRY(g->upload(&g->tile_type, tile_type));
  G_TRY(g->upload(&g->node_ptr, node_ptr)); G_TRY(g->upload(&g->node_src, node_src)); G_TRY(g->upload(&g->mask, hmask));
  G_TRY(g->upload(&g->type_rows, type_rows));
  const size_t nP = (size_t)n * P, n1 = (size_t)n + 1;
  G_TRY(g->alloc(&g->noise, nP)); G_TRY(g->alloc(&g->xt, nP));
  const int ntab = g->Gr > 0 ? 3 : 2;
  for (int t = 0; t < ntab; ++t) {
    G_TRY(g->alloc(&g->enc_z1[t], n1 * CCSP_HH)); G_TRY(g->alloc(&g->enc_a1[t], n1 * CCSP_HH));
    G_TRY(g->alloc(&g->enc_z2[t], n1 * CCSP_H)); G_TRY(g->alloc(&g->enc_e[t], n1 * CCSP_H));
    G_TRY(g->alloc(&g->enc_dz2[t], n1 * CCSP_H)); G_TRY(g->alloc(&g->enc_dz1[t], n1 * CCSP_HH));
    G_TRY(cudaMemsetAsync(g->enc_e[t], 0, n1 * CCSP_H * sizeof(float), st));      // row n stays zero: what padded rows gather
  }
  G_TRY(g->alloc(&g->emb, CCSP_H)); G_TRY(g->alloc(&g->tz1, 4 * CCSP_H)); G_TRY(g->alloc(&g->ta1, 4 * CCSP_H));
  G_TRY(g->alloc(&g->temb, CCSP_H)); G_TRY(g->alloc(&g->dtemb, CCSP_H));
  G_TRY(g->alloc(&g->bias, (size_t)MAX_TYPES * CCSP_H2)); G_TRY(g->alloc(&g->dbias, (size_t)MAX_TYPES * CCSP_H2));
  const size_t Ep = (size_t)Epad;
  G_TRY(g->alloc(&g->Z, Ep * CCSP_H2)); G_TRY(g->alloc(&g->H, Ep * CCSP_H2)); G_TRY(g->alloc(&g->dZ, Ep * CCSP_H2));
  G_TRY(g->alloc(&g->D1, 2 * Ep * CCSP_HH)); G_TRY(g->alloc(&g->A1, 2 * Ep * CCSP_HH)); G_TRY(g->alloc(&g->dD1, 2 * Ep * CCSP_HH));
  G_TRY(g->alloc(&g->O, 2 * Ep * P)); G_TRY(g->alloc(&g->dO, 2 * Ep * P));
  G_TRY(g->alloc(&g->out, nP)); G_TRY(g->alloc(&g->dOut, nP)); G_TRY(g->alloc(&g->err, (size_t)n));
  G_TRY(g->alloc(&g->dIn, Ep * g->Kseg));
  g->part_floats = (size_t)64 * CCSP_HH * CCSP_H;
  G_TRY(g->alloc(&g->part, g->part_floats));
  G_TRY(g->alloc(&g->colpart, (size_t)COLSUM_SLICES * MAX_TYPES * CCSP_H2));
  G_TRY(g->alloc(&g->dtemb_part, (size_t)MAX_TYPES * 8 * CCSP_H));
  G_TRY(cudaStreamSynchronize(st));
#undef G_TRY
  *out = g;
  return CCSP_OK;
}

void ccsp_train_graph_destroy(CcspTrainGraph *g) {
  if (!g) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(g->device);
  cudaDeviceSynchronize();
  g->free_all();
  delete g;
  cudaSetDevice(prev);
}

int64_t ccsp_train_graph_num_edges_of_type(const CcspTrainGraph *g, int32_t c) {
  return (g && c >= 0 && c < g->C) ? g->rows_of_type[c] : -1;
}

int ccsp_train_step(CcspTrainGraph *g, const CcspParams *w, const CcspParams *dw, int32_t t, float sqrt_alphas_cumprod_t,
                    float sqrt_one_minus_alphas_cumprod_t, const float *noise, int32_t loss_l1, float grad_scale, float *loss_out,
                    float *out_recon, void *stream) {
  TR_REQUIRE(g && w && dw && noise && loss_out, "null argument");
  TR_REQUIRE(t >= 0 && t < (1 << 20), "timestep out of range");
  TR_REQUIRE(w->geom_w0 && w->pose_w0 && w->dec_w0 && w->time_w1 && w->mlp_w && w->mlp_b, "null weight pointer");
  TR_REQUIRE(dw->geom_w0 && dw->pose_w0 && dw->dec_w0 && dw->time_w1 && dw->mlp_w && dw->mlp_b, "null gradient pointer");
  TR_REQUIRE(g->Gr == 0 || (w->grasp_w0 && dw->grasp_w0), "grasp encoder pointers missing");
  cudaStream_t st = (cudaStream_t)stream;
  CCSP_CUDA_TRY(cudaSetDevice(g->device));
  const int n = (int)g->n, P = g->P, C = g->C;
  const int64_t Epad = g->Epad;
  const int rows2 = (int)(2 * Epad);
  const int ntab = g->Gr > 0 ? 3 : 2;
  // tables: 0 geom, 1 pose, 2 grasp
  const EncPtrs ew[3] = {{w->geom_w0, w->geom_b0, w->geom_w2, w->geom_b2}, {w->pose_w0, w->pose_b0, w->pose_w2, w->pose_b2},
                         {w->grasp_w0, w->grasp_b0, w->grasp_w2, w->grasp_b2}};
  const EncGrads eg[3] = {{dw->geom_w0, dw->geom_b0, dw->geom_w2, dw->geom_b2}, {dw->pose_w0, dw->pose_b0, dw->pose_w2, dw->pose_b2},
                          {dw->grasp_w0, dw->grasp_b0, dw->grasp_w2, dw->grasp_b2}};
  const float *enc_in[3] = {g->x, g->xt, g->x};
  const int enc_ld[3] = {g->F, P, g->F}, enc_off[3] = {0, 0, g->grasp_begin}, enc_k[3] = {g->G, P, g->Gr};
  PtrTable W, B;
  MutPtrTable dW, dB;
  for (int c = 0; c < MAX_TYPES; ++c) {
    W.p[c] = c < C ? w->mlp_w[c] : nullptr; B.p[c] = c < C ? w->mlp_b[c] : nullptr;
    dW.p[c] = c < C ? dw->mlp_w[c] : nullptr; dB.p[c] = c < C ? dw->mlp_b[c] : nullptr;
    TR_REQUIRE(c >= C || (W.p[c] && B.p[c] && dW.p[c] && dB.p[c]), "null mlps pointer");
  }
  int rc;

  k_q_sample<<<(n * P + 255) / 256, 256, 0, st>>>(g->x0, noise, g->mask, n, P, sqrt_alphas_cumprod_t, sqrt_one_minus_alphas_cumprod_t,
                                                  g->noise, g->xt);
  CCSP_LAUNCH_CHECK();
  SegSrc in;
  if ((rc = forward_pass(g, w, W, B, t, in, st))) return rc;
  const float inv_count = 1.0f / (float)((size_t)n * P);
  k_node_loss<<<(n + 127) / 128, 128, 0, st>>>(g->O, g->node_ptr, g->node_src, g->mask, g->xtail, g->noise, n, P, g->normalize,
                                                loss_l1, inv_count, grad_scale, g->out, g->err, g->dOut);
  CCSP_LAUNCH_CHECK();
  k_loss_reduce<<<1, 256, 0, st>>>(g->err, n, inv_count, loss_out);
  CCSP_LAUNCH_CHECK();
  if (out_recon) CCSP_CUDA_TRY(cudaMemcpyAsync(out_recon, g->out, (size_t)n * P * sizeof(float), cudaMemcpyDeviceToDevice, st));

  // ---- backward --------------------------------------------------------------------------------------------------------
  if (Epad > 0) {
    k_dO<<<(rows2 * P + 255) / 256, 256, 0, st>>>(g->dOut, g->src_i, g->src_j, n, rows2, P, g->dO);
    CCSP_LAUNCH_CHECK();
    if ((rc = weight_grad(g, g->dO, g->A1, P, CCSP_HH, rows2, dw->dec_w2, st))) return rc;
    if ((rc = col_sum(g, g->dO, P, P, 0, rows2, dw->dec_b2, st))) return rc;
    k_dD1<<<(unsigned)(((size_t)rows2 * CCSP_HH + 255) / 256), 256, 0, st>>>(g->dO, w->dec_w2, g->D1, rows2, P, g->dD1);
    CCSP_LAUNCH_CHECK();
    if ((rc = weight_grad(g, g->dD1, g->H, CCSP_HH, CCSP_H, rows2, dw->dec_w0, st))) return rc;
    if ((rc = col_sum(g, g->dD1, CCSP_HH, CCSP_HH, 0, rows2, dw->dec_b0, st))) return rc;
    {
      LinearBwdInput p;
      p.dY = g->dD1; p.W = w->dec_w0; p.Zx = g->Z; p.dX = g->dZ; p.M_ = rows2; p.N_ = CCSP_H; p.K_ = CCSP_HH;
      if ((rc = launch_gemm(p, rows2, CCSP_H, 1, st))) return rc;
    }
    {  // db_c and, through them, the time columns
      ColSumArgs a;
      std::memset(&a, 0, sizeof(a));
      a.M = g->dZ; a.ld = CCSP_H2; a.cols = CCSP_H2; a.groups = C; a.out = g->dbias; a.part = g->colpart;
      for (int c = 0; c <= MAX_TYPES; ++c) a.start[c] = g->start[c];
      if ((rc = col_sums(a, st))) return rc;
      for (int c = 0; c < C; ++c)
        CCSP_CUDA_TRY(cudaMemcpyAsync(dB.p[c], g->dbias + (size_t)c * CCSP_H2, CCSP_H2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    {
      EdgeL1BwdWeight p;
      p.in = in; p.dZ = g->dZ; p.dW = dW; p.Kseg = g->Kseg; p.Kin = g->Kin;
      for (int c = 0; c <= MAX_TYPES; ++c) p.start[c] = g->start[c];
      if ((rc = launch_gemm(p, CCSP_H2, g->Kseg, C, st))) return rc;
    }
    {
      EdgeL1BwdInput p;
      p.dZ = g->dZ; p.W = W; p.tile_type = g->tile_type; p.dIn = g->dIn; p.Epad = (int)Epad; p.Kseg = g->Kseg; p.Kin = g->Kin; p.col0 = 0;
      if ((rc = launch_gemm(p, (int)Epad, g->Kseg, 1, st))) return rc;
    }
  } else {
    // no edges at all: every edge-side gradient is zero
    CCSP_CUDA_TRY(cudaMemsetAsync(dw->dec_w2, 0, (size_t)P * CCSP_HH * sizeof(float), st));
    CCSP_CUDA_TRY(cudaMemsetAsync(dw->dec_b2, 0, (size_t)P * sizeof(float), st));
    CCSP_CUDA_TRY(cudaMemsetAsync(dw->dec_w0, 0, (size_t)CCSP_HH * CCSP_H * sizeof(float), st));
    CCSP_CUDA_TRY(cudaMemsetAsync(dw->dec_b0, 0, (size_t)CCSP_HH * sizeof(float), st));
    CCSP_CUDA_TRY(cudaMemsetAsync(g->dbias, 0, (size_t)MAX_TYPES * CCSP_H2 * sizeof(float), st));
    for (int c = 0; c < C; ++c) {
      CCSP_CUDA_TRY(cudaMemsetAsync(dW.p[c], 0, (size_t)CCSP_H2 * g->Kin * sizeof(float), st));
      CCSP_CUDA_TRY(cudaMemsetAsync(dB.p[c], 0, (size_t)CCSP_H2 * sizeof(float), st));
    }
  }
  k_time_cols_bwd<<<dim3(C, 8), 256, 0, st>>>(W, dW, g->Kin, g->Kseg, g->temb, g->dbias, g->type_rows, g->dtemb_part);
  CCSP_LAUNCH_CHECK();
  k_time_bwd<<<4 * CCSP_H / 16, 256, 0, st>>>(g->dtemb_part, C * 8, g->emb, g->tz1, g->ta1, w->time_w3, dw->time_w1, dw->time_b1, dw->time_w3,
                                               dw->time_b3);
  CCSP_LAUNCH_CHECK();
  {
    NodeBwdArgs a;
    std::memset(&a, 0, sizeof(a));
    a.dIn = g->dIn; a.node_ptr = g->node_ptr; a.node_src = g->node_src; a.Kseg = g->Kseg; a.ntab = ntab;
    const int o = g->Gr > 0 ? 1 : 0;
    a.seg_of[0][0] = o; a.seg_of[0][1] = o + 1;             // geom_i, geom_j
    a.seg_of[1][0] = o + 2; a.seg_of[1][1] = o + 3;         // pose_i, pose_j
    a.seg_of[2][0] = 0; a.seg_of[2][1] = -1;                // grasp of arg1 only
    for (int tb = 0; tb < ntab; ++tb) { a.z2[tb] = g->enc_z2[tb]; a.dz2[tb] = g->enc_dz2[tb]; }
    k_node_bwd<<<n, 256, 0, st>>>(a);
    CCSP_LAUNCH_CHECK();
  }
  for (int tb = 0; tb < ntab; ++tb) {
    if ((rc = weight_grad(g, g->enc_dz2[tb], g->enc_a1[tb], CCSP_H, CCSP_HH, n, eg[tb].w2, st))) return rc;
    if ((rc = col_sum(g, g->enc_dz2[tb], CCSP_H, CCSP_H, 0, n, eg[tb].b2, st))) return rc;
    LinearBwdInput p;
    p.dY = g->enc_dz2[tb]; p.W = ew[tb].w2; p.Zx = g->enc_z1[tb]; p.dX = g->enc_dz1[tb]; p.M_ = n; p.N_ = CCSP_HH; p.K_ = CCSP_H;
    if ((rc = launch_gemm(p, n, CCSP_HH, 1, st))) return rc;
    k_enc1_bwd<<<CCSP_HH, 256, 0, st>>>(enc_in[tb], enc_ld[tb], enc_off[tb], enc_k[tb], n, g->enc_dz1[tb], eg[tb].w0, eg[tb].b0);
    CCSP_LAUNCH_CHECK();
  }
  return CCSP_OK;
}

int ccsp_energy_grad(CcspTrainGraph *g, const CcspParams *w, int32_t t, const float *x, float *energy_out, float *grad_out,
                     void *stream) {
  TR_REQUIRE(g && w && x && energy_out && grad_out, "null argument");
  TR_REQUIRE(t >= 0 && t < (1 << 20), "timestep out of range");
  TR_REQUIRE(w->geom_w0 && w->pose_w0 && w->dec_w0 && w->time_w1 && w->mlp_w && w->mlp_b, "null weight pointer");
  TR_REQUIRE(g->Gr == 0 || w->grasp_w0, "grasp encoder pointers missing");
  cudaStream_t st = (cudaStream_t)stream;
  CCSP_CUDA_TRY(cudaSetDevice(g->device));
  const int n = (int)g->n, P = g->P, C = g->C;
  const int64_t Epad = g->Epad;
  const int rows2 = (int)(2 * Epad);
  PtrTable W, B;
  for (int c = 0; c < MAX_TYPES; ++c) {
    W.p[c] = c < C ? w->mlp_w[c] : nullptr; B.p[c] = c < C ? w->mlp_b[c] : nullptr;
    TR_REQUIRE(c >= C || (W.p[c] && B.p[c]), "null mlps pointer");
  }
  int rc;
  CCSP_CUDA_TRY(cudaMemcpyAsync(g->xt, x, (size_t)n * P * sizeof(float), cudaMemcpyDeviceToDevice, st));
  SegSrc in;
  if ((rc = forward_pass(g, w, W, B, t, in, st))) return rc;
  if (Epad == 0) {
    CCSP_CUDA_TRY(cudaMemsetAsync(energy_out, 0, sizeof(float), st));
    CCSP_CUDA_TRY(cudaMemsetAsync(grad_out, 0, (size_t)n * P * sizeof(float), st));
    return CCSP_OK;
  }
  // E and dE/do, then back through decoder -> first layer (pose columns only) -> pose encoder, plus the direct -x term
  float *err = g->dIn;                       // dIn is [Epad, Kseg] >= 2 Epad floats; the error terms are consumed before it is rewritten
  k_energy_dO<<<(rows2 + 255) / 256, 256, 0, st>>>(g->O, g->xt, g->src_i, g->src_j, n, rows2, P, g->dO, err);
  CCSP_LAUNCH_CHECK();
  k_loss_reduce<<<1, 256, 0, st>>>(err, rows2, 1.0f, energy_out);
  CCSP_LAUNCH_CHECK();
  k_dD1<<<(unsigned)(((size_t)rows2 * CCSP_HH + 255) / 256), 256, 0, st>>>(g->dO, w->dec_w2, g->D1, rows2, P, g->dD1);
  CCSP_LAUNCH_CHECK();
  {
    LinearBwdInput p;
    p.dY = g->dD1; p.W = w->dec_w0; p.Zx = g->Z; p.dX = g->dZ; p.M_ = rows2; p.N_ = CCSP_H; p.K_ = CCSP_HH;
    if ((rc = launch_gemm(p, rows2, CCSP_H, 1, st))) return rc;
  }
  {
    EdgeL1BwdInput p;
    p.dZ = g->dZ; p.W = W; p.tile_type = g->tile_type; p.dIn = g->dIn; p.Epad = (int)Epad; p.Kseg = CCSP_H2; p.Kin = g->Kin;
    p.col0 = g->Kseg - CCSP_H2;              // the two pose segments are the last gathered ones (denoise_fn.py:346-354)
    if ((rc = launch_gemm(p, (int)Epad, CCSP_H2, 1, st))) return rc;
  }
  {
    NodeBwdArgs a;
    std::memset(&a, 0, sizeof(a));
    a.dIn = g->dIn; a.node_ptr = g->node_ptr; a.node_src = g->node_src; a.Kseg = CCSP_H2; a.ntab = 1;
    a.seg_of[0][0] = 0; a.seg_of[0][1] = 1;
    a.z2[0] = g->enc_z2[1]; a.dz2[0] = g->enc_dz2[1];
    k_node_bwd<<<n, 256, 0, st>>>(a);
    CCSP_LAUNCH_CHECK();
  }
  {
    LinearBwdInput p;
    p.dY = g->enc_dz2[1]; p.W = w->pose_w2; p.Zx = g->enc_z1[1]; p.dX = g->enc_dz1[1]; p.M_ = n; p.N_ = CCSP_HH; p.K_ = CCSP_H;
    if ((rc = launch_gemm(p, n, CCSP_HH, 1, st))) return rc;
  }
  k_energy_grad<<<(n * P + 255) / 256, 256, 0, st>>>(g->enc_dz1[1], w->pose_w0, g->dO, g->node_ptr, g->node_src, n, P, grad_out);
  CCSP_LAUNCH_CHECK();
  return CCSP_OK;
}

int ccsp_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t count, int32_t step, float lr,
                   float beta1, float beta2, float eps, void *stream) {
  TR_REQUIRE(param && grad && exp_avg && exp_avg_sq && count >= 0 && step >= 1, "bad Adam arguments");
  if (count == 0) return CCSP_OK;
  const float bc1 = 1.0f - (float)std::pow((double)beta1, (double)step);
  const float bc2 = 1.0f - (float)std::pow((double)beta2, (double)step);
  k_adam<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, (size_t)count, lr, beta1,
                                                                             beta2, eps, bc1, std::sqrt(bc2));
  CCSP_LAUNCH_CHECK();
  return CCSP_OK;
}

}  // extern "C"
