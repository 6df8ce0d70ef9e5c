// ccsp_abi.cu — host side of libccsp_b200.so: model packing, plan building, the sampling loop, and
// the extern "C" entry points declared in include/ccsp_b200.h.
#include "../../include/ccsp_b200.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels_simt.cuh"
#include "kernels_tc.cuh"
#include "kernels_fused2.cuh"
#include "kernels_node_tc.cuh"
#include "kernels_check.cuh"

namespace ccsp {
static thread_local std::string g_last_error;
static thread_local uint64_t g_launches = 0;
void set_error(const std::string &msg) { g_last_error = msg; }
void count_launch() { ++g_launches; }
}  // namespace ccsp

using namespace ccsp;

#define CCSP_REQUIRE(cond, msg)                      \
  do {                                               \
    if (!(cond)) {                                   \
      set_error(std::string("invalid argument: ") + msg); \
      return CCSP_ERR_INVALID;                       \
    }                                                \
  } while (0)

namespace {

// Device blocks of destroyed plans, kept by the model for the next plan: rebuilding a plan per batch (the end-to-end
// path) would otherwise pay cudaFree/cudaMalloc of ~0.5 GB every time, and cudaFree of large blocks was measured at
// 200-450 ms on the B200 boxes (scripts/plan_probe.py).
struct BlockCache {
  std::vector<std::pair<void *, size_t>> blocks;
  size_t bytes = 0;
  static constexpr size_t kCap = (size_t)8 << 30;
  void *take(size_t need, size_t *actual) {
    int best = -1;
    for (int i = 0; i < (int)blocks.size(); ++i)
      if (blocks[i].second >= need && blocks[i].second <= 2 * need + ((size_t)1 << 20) &&
          (best < 0 || blocks[i].second < blocks[best].second))
        best = i;
    if (best < 0) return nullptr;
    void *p = blocks[best].first;
    *actual = blocks[best].second;
    bytes -= blocks[best].second;
    blocks.erase(blocks.begin() + best);
    return p;
  }
  void give(void *p, size_t sz) {
    if (bytes + sz > kCap) { cudaFree(p); return; }
    blocks.emplace_back(p, sz);
    bytes += sz;
  }
  void clear() {
    for (auto &b : blocks) cudaFree(b.first);
    blocks.clear();
    bytes = 0;
  }
};

struct DevPool {   // owns device allocations of a model / plan
  std::vector<std::pair<void *, size_t>> ptrs;
  BlockCache *cache = nullptr;    // plans: the owning model's cache
  template <typename T>
  cudaError_t alloc(T **out, size_t count) {
    size_t bytes = ((count ? count : 1) * sizeof(T) + 255) & ~(size_t)255;
    size_t actual = bytes;        // a cached block can be larger than the request: keep its real size on record
    void *p = cache ? cache->take(bytes, &actual) : nullptr;
    cudaError_t e = cudaSuccess;
    if (!p) { actual = bytes; e = cudaMalloc(&p, bytes); }
    if (e == cudaSuccess) ptrs.emplace_back(p, actual);
    *out = (T *)p;
    return e;
  }
  void drop(std::pair<void *, size_t> &q) {
    if (!q.first) return;
    if (cache) cache->give(q.first, q.second);
    else cudaFree(q.first);
    q.first = nullptr;
  }
  void release(void *p) {
    for (auto &q : ptrs)
      if (q.first == p) drop(q);
  }
  void free_all() {
    for (auto &q : ptrs) drop(q);
    ptrs.clear();
  }
};

static thread_local int64_t g_upload_bytes = 0;

template <typename T>
cudaError_t upload(DevPool &pool, T **dst, const std::vector<T> &h) {
  cudaError_t e = pool.alloc(dst, h.size());
  if (e != cudaSuccess) return e;
  if (h.empty()) return cudaSuccess;
  g_upload_bytes += (int64_t)(h.size() * sizeof(T));
  return cudaMemcpy(*dst, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

// copy a reference-layout matrix (host OR device pointer) to a host vector
cudaError_t fetch(const float *src, size_t count, std::vector<float> &out) {
  out.resize(count);
  return cudaMemcpy(out.data(), src, count * sizeof(float), cudaMemcpyDefault);
}

// [rows, cols] row-major -> [cols, rows] row-major, restricted to columns [c0, c1)
std::vector<float> transpose_cols(const std::vector<float> &w, int rows, int cols, int c0, int c1) {
  std::vector<float> t((size_t)(c1 - c0) * rows);
  for (int r = 0; r < rows; ++r)
    for (int c = c0; c < c1; ++c) t[(size_t)(c - c0) * rows + r] = w[(size_t)r * cols + c];
  return t;
}

struct Encoder {
  float *w0 = nullptr, *b0 = nullptr, *w2t = nullptr, *b2 = nullptr;
  int din = 0;
};

}  // namespace

struct CcspPlan;
struct CcspModel {
  int device = 0;
  int G = 0, P = 0, Gr = 0, C = 0, normalize = 1, math = CCSP_MATH_FP32;
  int nseg_static = 2;            // 256-wide static segments of the first layer: [ (grasp_i) | geom_i | geom_j ]
  DevPool pool;
  Encoder geom, grasp, pose;
  float *dec_w1t = nullptr, *dec_b1 = nullptr, *dec_w2 = nullptr, *dec_b2 = nullptr;
  float *time_w1t = nullptr, *time_b1 = nullptr, *time_w3t = nullptr, *time_b3 = nullptr, *freqs = nullptr;
  float *Wst = nullptr;           // [C][Ks][512]  static columns, transposed
  float *Wpt = nullptr;           // [C][512][512] pose columns, transposed
  float *Wtt = nullptr;           // [C][256][512] time columns, transposed
  float *bias = nullptr;          // [C][512]
  float *tb = nullptr;            // [tb_T][C][512] per-(t, type) time term
  int tb_T = 0;
  // tensor-core operand blobs (kernels_tc.cuh), packed lazily per CcspMath mode
  int num_sms = 148;
  std::vector<float> h_pose_w;    // [C][512][512] pose columns of mlps[c].weight, reference [out][k] layout
  std::vector<float> h_dec_w1;    // [128][256]
  std::vector<float> h_pose_w2;   // [256][128] pose_encoder.2.weight (B operand of the tcgen05 node kernel)
  uint8_t *blob_pose[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  BlockCache cache;               // device blocks handed back by destroyed plans
  std::vector<CcspPlan *> plans;  // live plans (detached from the cache if the model is destroyed first)
  uint8_t *blob_l1[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  uint8_t *blob_dec[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
};

struct CcspPlan {
  CcspModel *m = nullptr;
  int device = 0;                 // the model's device, kept so that destroy works after the model is gone
  int64_t n = 0, E = 0, Epad = 0;
  int num_tiles = 0;
  int num_chains = 1;             // independent scene groups; rows are grouped (chain, type)
  int chain_tile0[CCSP_MAX_CHAINS + 1] = {0};   // tiles of chain c: [chain_tile0[c], chain_tile0[c + 1])
  int chain_row0[CCSP_MAX_CHAINS + 1] = {0};    // nodes of chain c: [chain_row0[c], chain_row0[c + 1])
  DevPool pool;
  int *src_i = nullptr, *src_j = nullptr, *tile_type = nullptr, *node_ptr = nullptr, *node_src = nullptr;
  signed char *mask = nullptr;
  float *gt = nullptr, *xtail = nullptr;
  float *S = nullptr;             // [Epad, 512] static pre-activation, blocked layout (common.cuh blk_off)
  float *o = nullptr;             // [Epad, 2, P] decoder outputs
  float *x = nullptr;             // [n, P] sampler state
  // per-arithmetic-mode work buffers, allocated on first use (ensure_mode_buffers)
  float *pe32 = nullptr;          // FP32 path: [n+1, 256] pose embeddings (row n = zeros for padded edges)
  float *H32 = nullptr;           // FP32 path: [Epad, 512] first-layer activations, row-major
  uint8_t *pe_split[2] = {nullptr, nullptr};   // tensor-core paths, per operand kind (TF32 / BF16): [n+1][256 hi | 256 lo]
  uint8_t *Hop[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // per math mode: H in decoder-operand format
  int64_t h2d_bytes = 0;
  // sampled kernel timing (bench roofline)
  int timing_stride = 0;
  std::vector<cudaEvent_t> ev;    // groups of 4: before l1 | after l1 | after dec | after node
  size_t ev_used = 0;
  CcspTiming timing = {0, 0.0, 0.0, 0.0};
  cudaStream_t capture_stream = nullptr;
  // persistent small-shard path (sample_persistent): two side streams, their events, the flag pair, device schedule arrays
  cudaStream_t ps_node = nullptr, ps_edge = nullptr;
  cudaEvent_t pe_begin = nullptr, pe_node = nullptr, pe_edge = nullptr, pe_t0 = nullptr, pe_t1 = nullptr;
  unsigned *arrive_host = nullptr, *arrive_dev = nullptr;   // host-mapped: CTAs of the persistent edge kernel that have started
  int persist_timed_evals = 0;    // > 0: pe_t0 / pe_t1 bracket a persistent edge kernel that ran this many evaluations
  unsigned *p_flags = nullptr;
  NodeEval *p_sched = nullptr;
  int *p_eval_t = nullptr;
  size_t p_sched_cap = 0;
  cudaGraphExec_t graph_exec = nullptr;   // CCSP_GRAPH=1: the last captured sampling loop (kept alive until the next call / destroy)
};

// -------------------------------------------------------------------------------------------------
static int upload_encoder(CcspModel *m, Encoder &enc, const float *w0, const float *b0, const float *w2,
                          const float *b2, int din) {
  std::vector<float> h;
  enc.din = din;
  CCSP_CUDA_TRY(fetch(w0, (size_t)CCSP_HH * din, h));
  CCSP_CUDA_TRY(upload(m->pool, &enc.w0, h));
  CCSP_CUDA_TRY(fetch(b0, CCSP_HH, h));
  CCSP_CUDA_TRY(upload(m->pool, &enc.b0, h));
  CCSP_CUDA_TRY(fetch(w2, (size_t)CCSP_H * CCSP_HH, h));
  CCSP_CUDA_TRY(upload(m->pool, &enc.w2t, transpose_cols(h, CCSP_H, CCSP_HH, 0, CCSP_HH)));
  CCSP_CUDA_TRY(fetch(b2, CCSP_H, h));
  CCSP_CUDA_TRY(upload(m->pool, &enc.b2, h));
  return CCSP_OK;
}

static int model_build(CcspModel *m, const CcspModelDesc *d) {
  CCSP_CUDA_TRY(cudaGetDevice(&m->device));
  m->G = d->geom_dim; m->P = d->pose_dim; m->Gr = d->grasp_dim; m->C = d->num_types;
  m->normalize = d->normalize ? 1 : 0;
  m->nseg_static = d->grasp_dim > 0 ? 3 : 2;
  int rc;
  if ((rc = upload_encoder(m, m->geom, d->geom_w0, d->geom_b0, d->geom_w2, d->geom_b2, m->G))) return rc;
  if (m->Gr > 0)
    if ((rc = upload_encoder(m, m->grasp, d->grasp_w0, d->grasp_b0, d->grasp_w2, d->grasp_b2, m->Gr))) return rc;
  if ((rc = upload_encoder(m, m->pose, d->pose_w0, d->pose_b0, d->pose_w2, d->pose_b2, m->P))) return rc;
  CCSP_CUDA_TRY(fetch(d->pose_w2, (size_t)CCSP_H * CCSP_HH, m->h_pose_w2));

  std::vector<float> h;
  CCSP_CUDA_TRY(fetch(d->dec_w0, (size_t)CCSP_HH * CCSP_H, h));
  m->h_dec_w1 = h;
  CCSP_CUDA_TRY(upload(m->pool, &m->dec_w1t, transpose_cols(h, CCSP_HH, CCSP_H, 0, CCSP_H)));
  CCSP_CUDA_TRY(fetch(d->dec_b0, CCSP_HH, h));
  CCSP_CUDA_TRY(upload(m->pool, &m->dec_b1, h));
  CCSP_CUDA_TRY(fetch(d->dec_w2, (size_t)m->P * CCSP_HH, h));
  CCSP_CUDA_TRY(upload(m->pool, &m->dec_w2, h));
  CCSP_CUDA_TRY(fetch(d->dec_b2, m->P, h));
  CCSP_CUDA_TRY(upload(m->pool, &m->dec_b2, h));

  CCSP_CUDA_TRY(fetch(d->time_w1, (size_t)4 * CCSP_H * CCSP_H, h));
  CCSP_CUDA_TRY(upload(m->pool, &m->time_w1t, transpose_cols(h, 4 * CCSP_H, CCSP_H, 0, CCSP_H)));
  CCSP_CUDA_TRY(fetch(d->time_b1, 4 * CCSP_H, h));
  CCSP_CUDA_TRY(upload(m->pool, &m->time_b1, h));
  CCSP_CUDA_TRY(fetch(d->time_w3, (size_t)CCSP_H * 4 * CCSP_H, h));
  CCSP_CUDA_TRY(upload(m->pool, &m->time_w3t, transpose_cols(h, CCSP_H, 4 * CCSP_H, 0, 4 * CCSP_H)));
  CCSP_CUDA_TRY(fetch(d->time_b3, CCSP_H, h));
  CCSP_CUDA_TRY(upload(m->pool, &m->time_b3, h));
  {  // SinusoidalPosEmb frequencies, FP32 like torch.exp(torch.arange(128) * -emb)  (denoise_fn.py:45-47)
    std::vector<float> f(CCSP_HH);
    const float e = (float)(-(std::log(10000.0) / (CCSP_HH - 1)));
    for (int k = 0; k < CCSP_HH; ++k) f[k] = expf((float)k * e);
    CCSP_CUDA_TRY(upload(m->pool, &m->freqs, f));
  }

  const int Ks = m->nseg_static * CCSP_H, Kin = Ks + 3 * CCSP_H;
  std::vector<float> wst((size_t)m->C * Ks * CCSP_H2), wpt((size_t)m->C * CCSP_H2 * CCSP_H2),
      wtt((size_t)m->C * CCSP_H * CCSP_H2), bias((size_t)m->C * CCSP_H2);
  m->h_pose_w.resize((size_t)m->C * CCSP_H2 * CCSP_H2);
  {
    cudaDeviceProp prop;
    CCSP_CUDA_TRY(cudaGetDeviceProperties(&prop, m->device));
    m->num_sms = prop.multiProcessorCount;
  }
  for (int c = 0; c < m->C; ++c) {
    CCSP_CUDA_TRY(fetch(d->mlp_w[c], (size_t)CCSP_H2 * Kin, h));
    for (int r = 0; r < CCSP_H2; ++r)
      std::memcpy(&m->h_pose_w[((size_t)c * CCSP_H2 + r) * CCSP_H2], &h[(size_t)r * Kin + Ks], CCSP_H2 * sizeof(float));
    auto a = transpose_cols(h, CCSP_H2, Kin, 0, Ks);
    auto b = transpose_cols(h, CCSP_H2, Kin, Ks, Ks + CCSP_H2);
    auto t = transpose_cols(h, CCSP_H2, Kin, Ks + CCSP_H2, Kin);
    std::memcpy(&wst[(size_t)c * Ks * CCSP_H2], a.data(), a.size() * sizeof(float));
    std::memcpy(&wpt[(size_t)c * CCSP_H2 * CCSP_H2], b.data(), b.size() * sizeof(float));
    std::memcpy(&wtt[(size_t)c * CCSP_H * CCSP_H2], t.data(), t.size() * sizeof(float));
    CCSP_CUDA_TRY(fetch(d->mlp_b[c], CCSP_H2, h));
    std::memcpy(&bias[(size_t)c * CCSP_H2], h.data(), CCSP_H2 * sizeof(float));
  }
  CCSP_CUDA_TRY(upload(m->pool, &m->Wst, wst));
  CCSP_CUDA_TRY(upload(m->pool, &m->Wpt, wpt));
  CCSP_CUDA_TRY(upload(m->pool, &m->Wtt, wtt));
  CCSP_CUDA_TRY(upload(m->pool, &m->bias, bias));
  return CCSP_OK;
}

// time-term table tb[t][c][:] for t < T  (run-constant: depends on the weights and t only)
static int ensure_time_table(CcspModel *m, int T, cudaStream_t st) {
  if (T <= m->tb_T) return CCSP_OK;
  if (m->tb) { CCSP_CUDA_TRY(cudaDeviceSynchronize()); m->pool.release(m->tb);   /* other streams may still read the old table */ m->tb = nullptr; m->tb_T = 0; }
  float *temb = nullptr;
  CCSP_CUDA_TRY(cudaMalloc(&temb, (size_t)T * CCSP_H * sizeof(float)));
  CCSP_CUDA_TRY(m->pool.alloc(&m->tb, (size_t)T * m->C * CCSP_H2));
  k_time_embed<<<T, 256, 0, st>>>(m->freqs, m->time_w1t, m->time_b1, m->time_w3t, m->time_b3, temb);
  CCSP_LAUNCH_CHECK();
  k_time_bias<<<dim3(T, m->C), 256, 0, st>>>(temb, m->Wtt, m->C, m->tb);
  CCSP_LAUNCH_CHECK();
  CCSP_CUDA_TRY(cudaStreamSynchronize(st));
  CCSP_CUDA_TRY(cudaFree(temb));
  m->tb_T = T;
  return CCSP_OK;
}

// ---- tensor-core modes -----------------------------------------------------------------------------
template <class M>
static int pack_tc_blobs(CcspModel *m, int math) {
  using L1 = tc::L1Cfg<M, 1>;
  using Dec = tc::DecCfg<M>;
  const size_t per = (size_t)2 * M::NKC1 * L1::B_STAGE;
  std::vector<uint8_t> b1(per * m->C);
  for (int c = 0; c < m->C; ++c)
    tc::pack_b_blob<M, 256>(&m->h_pose_w[(size_t)c * CCSP_H2 * CCSP_H2], CCSP_H2, 0, CCSP_H2, CCSP_H2, b1.data() + c * per);
  std::vector<uint8_t> b2((size_t)M::NKC2 * Dec::B_STAGE);
  tc::pack_b_blob<M, 128>(m->h_dec_w1.data(), CCSP_H, 0, CCSP_H, CCSP_HH, b2.data());
  CCSP_CUDA_TRY(upload(m->pool, &m->blob_l1[math], b1));
  CCSP_CUDA_TRY(upload(m->pool, &m->blob_dec[math], b2));
  if (M::KIND == tc::KIND_BF16) {
    std::vector<uint8_t> b3((size_t)tc::NodeTcCfg<tc::Mode<tc::KIND_BF16, M::NSPLIT>>::NKC * M::NS * CCSP_H * tc::ROWB);
    tc::pack_b_blob<M, CCSP_H>(m->h_pose_w2.data(), CCSP_HH, 0, CCSP_HH, CCSP_H, b3.data());
    CCSP_CUDA_TRY(upload(m->pool, &m->blob_pose[math], b3));
  }
  return CCSP_OK;
}

static int ensure_tc_blobs(CcspModel *m, int math) {
  if (math == CCSP_MATH_FP32 || m->blob_l1[math]) return CCSP_OK;
  switch (math) {
    case CCSP_MATH_TF32X3: return pack_tc_blobs<tc::Mode<tc::KIND_TF32, 3>>(m, math);
    case CCSP_MATH_BF16X3: return pack_tc_blobs<tc::Mode<tc::KIND_BF16, 3>>(m, math);
    case CCSP_MATH_TF32: return pack_tc_blobs<tc::Mode<tc::KIND_TF32, 1>>(m, math);
    case CCSP_MATH_BF16: return pack_tc_blobs<tc::Mode<tc::KIND_BF16, 1>>(m, math);
  }
  set_error("unknown math mode");
  return CCSP_ERR_INVALID;
}

static inline int math_kind(int math) { return (math == CCSP_MATH_TF32X3 || math == CCSP_MATH_TF32) ? 0 : 1; }
static inline int math_ns(int math) { return (math == CCSP_MATH_TF32X3 || math == CCSP_MATH_BF16X3) ? 2 : 1; }

// work buffers of the plan for the model's current arithmetic mode
static int ensure_mode_buffers(CcspPlan *p) {
  CcspModel *m = p->m;
  if (m->math == CCSP_MATH_FP32) {
    if (!p->pe32) {
      CCSP_CUDA_TRY(p->pool.alloc(&p->pe32, (size_t)(p->n + 1) * CCSP_H));
      CCSP_CUDA_TRY(p->pool.alloc(&p->H32, (size_t)p->Epad * CCSP_H2));
    }
    return CCSP_OK;
  }
  const int kind = math_kind(m->math), elt = kind == 0 ? 4 : 2;
  if (!p->pe_split[kind]) CCSP_CUDA_TRY(p->pool.alloc(&p->pe_split[kind], (size_t)(p->n + 1) * 2 * CCSP_H * elt));
  // BF16 operand modes run the fused CTA-pair kernel (H never leaves the SM); only the two-kernel TF32 path needs H in HBM
  if (kind == 0 && !p->Hop[m->math]) CCSP_CUDA_TRY(p->pool.alloc(&p->Hop[m->math], (size_t)p->Epad * CCSP_H2 * elt * math_ns(m->math)));
  return CCSP_OK;
}

static inline bool edge_fused(int math) { return math == CCSP_MATH_BF16X3 || math == CCSP_MATH_BF16; }

// BF16 operand modes: first layer + decoder as ONE CTA-pair kernel (kernels_fused2.cuh)
template <class M>
static int launch_edge_pair(CcspPlan *p, const float *tb, cudaStream_t st) {
  CcspModel *m = p->m;
  tc::FusedArgs a;
  std::memset(&a, 0, sizeof(a));
  a.pe_split = p->pe_split[1];
  a.src_i = p->src_i; a.src_j = p->src_j;
  a.b_blob = m->blob_l1[m->math]; a.w_blob = m->blob_dec[m->math];
  a.tile_type = p->tile_type;
  a.num_m_tiles = (int)(p->Epad / CCSP_TILE_M);
  a.S = p->S; a.tb = tb;
  a.bd1 = m->dec_b1; a.Wd2 = m->dec_w2; a.bd2 = m->dec_b2; a.P = m->P; a.o = p->o;
  a.pad_row_plus1 = (int)p->n + 1;
  CCSP_CUDA_TRY((tc::launch_fused2_tc<M>(a, m->num_sms, st)));
  count_launch();
  return CCSP_OK;
}

template <class M>
static int launch_edge_tc(CcspPlan *p, const float *tb, cudaStream_t st, cudaEvent_t mid) {
  CcspModel *m = p->m;
  tc::L1Args a;
  std::memset(&a, 0, sizeof(a));
  a.pe_split = p->pe_split[math_kind(m->math)];
  a.src_i = p->src_i; a.src_j = p->src_j;
  a.b_blob = m->blob_l1[m->math]; a.tile_type = p->tile_type;
  a.num_m_tiles = (int)(p->Epad / CCSP_TILE_M);
  a.S = p->S; a.tb = tb; a.H = p->Hop[m->math];
  CCSP_CUDA_TRY((tc::launch_l1_tc<M, CCSP_CLUSTER>(a, m->num_sms, st)));
  count_launch();
  if (mid) CCSP_CUDA_TRY(cudaEventRecord(mid, st));
  tc::DecArgs d;
  std::memset(&d, 0, sizeof(d));
  d.H = p->Hop[m->math];
  d.b_blob = m->blob_dec[m->math];
  d.num_tiles = (int)(2 * p->Epad / CCSP_TILE_M);
  d.bd1 = m->dec_b1; d.Wd2 = m->dec_w2; d.bd2 = m->dec_b2; d.P = m->P; d.o = p->o;
  CCSP_CUDA_TRY((tc::launch_dec_tc<M>(d, m->num_sms, st)));
  count_launch();
  return CCSP_OK;
}

// one evaluation of the two dense per-edge layers for timestep t: pe -> H -> o
static int launch_edge(CcspPlan *p, int t, cudaStream_t st, cudaEvent_t mid = nullptr) {
  CcspModel *m = p->m;
  if (p->Epad == 0) return CCSP_OK;
  const float *tb = m->tb + (size_t)t * m->C * CCSP_H2;
  switch (m->math) {
    case CCSP_MATH_FP32: {
      RowSrc rs;
      rs.nseg = 2;
      rs.src[0] = p->pe32; rs.idx[0] = p->src_i;
      rs.src[1] = p->pe32; rs.idx[1] = p->src_j;
      rs.src[2] = nullptr; rs.idx[2] = nullptr;
      k_edge_l1_simt<EPI_L1><<<dim3((unsigned)(p->Epad / SG_BM), CCSP_H2 / SG_BN), 256, 0, st>>>(
          rs, m->Wpt, p->tile_type, m->bias, p->S, tb, p->H32);
      CCSP_LAUNCH_CHECK();
      if (mid) CCSP_CUDA_TRY(cudaEventRecord(mid, st));
      k_edge_dec_simt<<<(unsigned)(2 * p->Epad / SG_BM), 256, 0, st>>>(p->H32, m->dec_w1t, m->dec_b1, m->dec_w2,
                                                                       m->dec_b2, m->P, p->o);
      CCSP_LAUNCH_CHECK();
      return CCSP_OK;
    }
    case CCSP_MATH_TF32X3: return launch_edge_tc<tc::Mode<tc::KIND_TF32, 3>>(p, tb, st, mid);
    case CCSP_MATH_BF16X3: return launch_edge_pair<tc::Mode<tc::KIND_BF16, 3>>(p, tb, st);
    case CCSP_MATH_TF32: return launch_edge_tc<tc::Mode<tc::KIND_TF32, 1>>(p, tb, st, mid);
    case CCSP_MATH_BF16: return launch_edge_pair<tc::Mode<tc::KIND_BF16, 1>>(p, tb, st);
    default:
      set_error("unknown math mode");
      return CCSP_ERR_STATE;
  }
}

static NodeArgs node_args_base(CcspPlan *p) {
  CcspModel *m = p->m;
  NodeArgs a;
  std::memset(&a, 0, sizeof(a));
  a.n = (int)p->n; a.P = m->P; a.normalize = m->normalize;
  a.x = p->x; a.o = p->o; a.node_ptr = p->node_ptr; a.node_src = p->node_src;
  a.mask = p->mask; a.gt = p->gt; a.xtail = p->xtail;
  a.W0 = m->pose.w0; a.b0 = m->pose.b0; a.W2t = m->pose.w2t; a.b2 = m->pose.b2;
  if (m->math == CCSP_MATH_FP32) {
    a.pe_fmt = 0; a.pe = p->pe32;
  } else {
    a.pe_fmt = 1 + math_kind(m->math); a.pe = p->pe_split[math_kind(m->math)];
  }
  return a;
}

static int launch_node(CcspPlan *p, const NodeArgs &a, cudaStream_t st) {
  static bool configured_dev[64] = {};      // per device: function attributes belong to the device's context
  int dev_ = 0;
  cudaGetDevice(&dev_);
  bool &configured = configured_dev[dev_ & 63];
  if (!configured) {
    CCSP_CUDA_TRY(cudaFuncSetAttribute(k_node, cudaFuncAttributeMaxDynamicSharedMemorySize, NODE_SMEM_BYTES));
    configured = true;
  }
  // BF16 operand modes: the pose encoder's 128 -> 256 layer runs on tcgen05 (kernels_node_tc.cuh)
  const int math = p->m->math;
  if (edge_fused(math) && a.mode != NODE_EPS_OUT) {
    static const bool want_trace = getenv("CCSP_NODE_TRACE") != nullptr;
    static long ncall = 0;
    if (want_trace && ++ncall == 500) {      // developer aid: phase timeline of one launch (CTA 1: thread 0 and the MMA thread)
      long long *d = nullptr, h[32];
      CCSP_CUDA_TRY(cudaMalloc(&d, sizeof(h)));
      CCSP_CUDA_TRY(cudaMemset(d, 0, sizeof(h)));
      NodeArgs b = a;
      b.trace = d;
      CCSP_CUDA_TRY(cudaStreamSynchronize(st));
      if (math == CCSP_MATH_BF16X3) CCSP_CUDA_TRY((tc::launch_node_tc<tc::Mode<tc::KIND_BF16, 3>>(b, p->m->blob_pose[math], st)));
      else CCSP_CUDA_TRY((tc::launch_node_tc<tc::Mode<tc::KIND_BF16, 1>>(b, p->m->blob_pose[math], st)));
      CCSP_CUDA_TRY(cudaStreamSynchronize(st));
      CCSP_CUDA_TRY(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
      cudaFree(d);
      fprintf(stderr, "[node trace] mode %d  t0:", a.mode);
      for (int i = 0; i < 16; ++i) fprintf(stderr, " %lld", h[i] ? h[i] - h[0] : -1);
      fprintf(stderr, "\n[node trace] mma thread:");
      for (int i = 0; i < 12; ++i) fprintf(stderr, " %lld", h[16 + i] ? h[16 + i] - h[0] : -1);
      fprintf(stderr, "\n");
      count_launch();
      return CCSP_OK;
    }
    if (math == CCSP_MATH_BF16X3) CCSP_CUDA_TRY((tc::launch_node_tc<tc::Mode<tc::KIND_BF16, 3>>(a, p->m->blob_pose[math], st)));
    else CCSP_CUDA_TRY((tc::launch_node_tc<tc::Mode<tc::KIND_BF16, 1>>(a, p->m->blob_pose[math], st)));
    count_launch();
    return CCSP_OK;
  }
  unsigned blocks = (unsigned)((p->n + 1 + NODE_ROWS - 1) / NODE_ROWS);
  k_node<<<blocks, NODE_THREADS, NODE_SMEM_BYTES, st>>>(a);
  CCSP_LAUNCH_CHECK();
  return CCSP_OK;
}


// host-mapped trap word (see kernels_tc.cuh::trap_with)
static unsigned long long *g_trap_host = nullptr;
static unsigned long long *g_ptrace_dev = nullptr;      // persistent-mode timeline: [8 events][32 iterations] ns (words 8.. of the block)
static void ensure_trap_word() {
  if (g_trap_host) return;
  if (cudaHostAlloc((void **)&g_trap_host, 8 * (8 + 12 * 32), cudaHostAllocMapped) != cudaSuccess) { g_trap_host = nullptr; return; }
  std::memset(g_trap_host, 0, 8 * (8 + 12 * 32));
  unsigned long long *dptr = nullptr;
  if (cudaHostGetDevicePointer((void **)&dptr, g_trap_host, 0) == cudaSuccess) {
    cudaMemcpyToSymbol(tc::g_trap_info, &dptr, sizeof(dptr));
    g_ptrace_dev = dptr + 8;
  }
}

// host-mapped arrival counter of a plan's persistent edge kernel (one increment per CTA at entry); owned by the plan
static void ensure_arrival_word(CcspPlan *p) {
  if (p->arrive_host) return;
  if (cudaHostAlloc((void **)&p->arrive_host, 64, cudaHostAllocMapped) != cudaSuccess) { p->arrive_host = nullptr; return; }
  *p->arrive_host = 0;
  if (cudaHostGetDevicePointer((void **)&p->arrive_dev, p->arrive_host, 0) != cudaSuccess) {
    cudaFreeHost(p->arrive_host);
    p->arrive_host = p->arrive_dev = nullptr;
  }
}

// -------------------------------------------------------------------------------------------------------------------------
// Persistent small-shard path.  When the whole sample fits on the chip at once — every edge unit on its own CTA pair plus all
// node CTAs, counted in TPCs so that cluster placement cannot be starved: pairs + node_ctas <= num_sms / 2 — the T x (1 + K)
// loop runs as TWO kernels launched ONCE (k_edge_fused2_tc<PERSIST>, k_node_tc<PERSIST>) on two side streams.  They hand-shake
// per evaluation through two device counters (release / acquire at gpu scope) instead of 2 launches per evaluation, so barrier
// and TMEM set-up, table loads and the W2 fetch are paid once per sample and the launch gaps disappear.  Same arithmetic in the
// same order: results are bit-identical to the launch-per-evaluation path (tests/test_gpu_persistent.py).
// -------------------------------------------------------------------------------------------------------------------------
// Cut the node range [0, n) into `want` contiguous groups of whole scenes: a cut between nodes s - 1 and s is legal when no
// (typed) edge has one endpoint on either side; among the legal cuts the k-th is the one whose edge count below it is closest
// to k / want of all edges.  Returns the group boundaries {0, .., n}; {0, n} when fewer legal cuts exist than asked for.
static std::vector<int64_t> chain_cuts(int64_t n, int64_t E, const int64_t *edge_index, const std::vector<int> &etype, int want) {
  std::vector<int64_t> whole = {0, n};
  if (want <= 1 || n < 2) return whole;
  std::vector<int> cover(n + 2, 0);
  std::vector<int64_t> below(n + 2, 0);                     // below[s] = edges with both endpoints < s
  int64_t Ev = 0;
  for (int64_t e = 0; e < E; ++e) {
    if (etype[e] < 0) continue;
    const int64_t i = edge_index[e], j = edge_index[E + e];
    const int64_t lo = std::min(i, j), hi = std::max(i, j);
    ++cover[lo + 1]; --cover[hi + 1];
    ++below[hi + 1];
    ++Ev;
  }
  if (Ev == 0) return whole;
  for (int64_t s2 = 1; s2 <= n; ++s2) { cover[s2] += cover[s2 - 1]; below[s2] += below[s2 - 1]; }
  std::vector<int64_t> out = {0};
  int64_t prev = 0;
  for (int k = 1; k < want; ++k) {
    const int64_t target = Ev * k / want;
    int64_t best = -1, best_d = -1;
    for (int64_t s2 = prev + 1; s2 < n; ++s2) {
      if (cover[s2] != 0) continue;                         // an edge spans the cut between nodes s2 - 1 and s2
      const int64_t d = std::llabs(below[s2] - target);
      if (best < 0 || d < best_d) { best = s2; best_d = d; }
      if (below[s2] > target) break;
    }
    if (best < 0) return whole;
    out.push_back(best);
    prev = best;
  }
  out.push_back(n);
  return out;
}

extern "C" { static int drain_timing(CcspPlan *p); }
struct PersistCfg {
  int pairs = 0;        // CTA pairs of the persistent edge kernel
  int node_ctas = 0;    // CTAs of the persistent node kernel
  int partition = 0;    // chains: 1 = every node CTA serves one chain, one block each (small shards)
};
static bool persistent_eligible(const CcspPlan *p, PersistCfg *cfg) {
  const CcspModel *m = p->m;
  const char *env = getenv("CCSP_PERSIST");
  if (env && env[0] == '0') return false;
  if (m->math != CCSP_MATH_BF16X3 || p->Epad == 0) return false;
  const int node_blocks = (int)((p->n + 1 + 63) / 64);
  const int units = (int)(p->Epad / CCSP_TILE_M);
  const int tpcs = m->num_sms / 2;
  if (p->num_chains > 1) {
    // Pipelined chains (plans cut into chains at creation, CCSP_CHAINS=2..4): a few node CTAs serve all node blocks of one
    // chain while the edge kernel, on all other SMs, works on the other chain.
    // A tool that serialises kernels (ncu, compute-sanitizer inject through these variables) would dead-lock two
    // co-operating kernels: stay on the launch-per-evaluation path there.
    if (getenv("CUDA_INJECTION64_PATH") || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || getenv("NSIGHT_CUDA_DEBUGGER")) return false;
    int bmax = 0;
    for (int c = 0; c < p->num_chains; ++c) bmax = std::max(bmax, (p->chain_row0[c + 1] - p->chain_row0[c] + 63) / 64);
    int nc = 24;
    if (bmax * p->num_chains <= 48 && !getenv("CCSP_PIPE_NODE_CTAS")) {
      // small shard: SMs are plentiful, so every chain gets its own node CTAs, one 64-node block each
      nc = bmax * p->num_chains;
      cfg->partition = 1;
    } else {
      if (const char *e2 = getenv("CCSP_PIPE_NODE_CTAS")) nc = atoi(e2);
      nc = std::max(2, std::min(nc, std::min(node_blocks, m->num_sms - 4)));
    }
    nc = (nc + 1) & ~1;                         // whole TPCs, so that the edge clusters keep whole TPCs too
    int pairs = (m->num_sms - nc) / 2;
    pairs = std::min(pairs, units);
    if (pairs < 1) return false;
    cfg->pairs = pairs; cfg->node_ctas = nc;
    return true;
  }
  if (p->timing_stride > 0) return false;
  const int pairs = std::min(units, tpcs - node_blocks);
  // single chain: only while the shard is latency-bound: at most two rounds of units per pair (beyond that the
  // launch-per-evaluation path amortises its fixed cost and uses all SMs for the edge phase)
  if (pairs < 1 || (units + pairs - 1) / pairs > 2) return false;
  cfg->pairs = pairs; cfg->node_ctas = node_blocks;
  // Opt-in only (CCSP_PERSIST=1).  Measured on B200 (profiles/README.md R2.6): once padded rows stopped hammering the zero row
  // (cp.async zero-fill), the single-chain persistent pair is no faster than two launches per evaluation on any shard size — the
  // per-evaluation floor is the dependent chain inside the phases, not launch overhead.
  return env && env[0] == '1';
}

static int sample_persistent(CcspPlan *p, const CcspSchedule *s, const CcspNoise *nz, float *out, float *history, cudaStream_t user_st,
                             const PersistCfg &cfg) {
  const int pairs = cfg.pairs;
  using M = tc::Mode<tc::KIND_BF16, 3>;
  CcspModel *m = p->m;
  const int T = s->T;
  const int per = s->ebm_per_steps > 0 ? s->ebm_per_steps : 1;
  const size_t nP = (size_t)p->n * m->P;
  // ---- the schedule of node iterations and edge evaluations, exactly the launch sequence of ccsp_sample -----------------
  std::vector<NodeEval> sched;
  std::vector<int> eval_t;
  sched.reserve((size_t)T * 11 + 1);
  unsigned draw = 0;
  {
    NodeEval e;
    std::memset(&e, 0, sizeof(e));
    e.mode = NODE_INIT; e.pin = 1; e.draw = draw++; e.hist_slot = history ? 0 : -1;
    sched.push_back(e);
  }
  for (int j = T - 1; j >= 0; --j) {
    const int Kt = (s->samples_per_step && (j % per == 0)) ? s->samples_per_step[j] : 0;
    const int hslot = history ? T - j : -1;
    NodeEval e;
    std::memset(&e, 0, sizeof(e));
    e.mode = NODE_DDPM;
    e.a = s->sqrt_recip_alphas_cumprod[j]; e.b = s->sqrt_recipm1_alphas_cumprod[j];
    e.c1 = s->posterior_mean_coef1[j]; e.c2 = s->posterior_mean_coef2[j];
    e.sigma = (j == 0 ? 0.f : 1.f) * expf(0.5f * s->posterior_log_variance_clipped[j]);
    e.pin = Kt == 0; e.hist_slot = Kt == 0 ? hslot : -1; e.draw = draw++;
    sched.push_back(e); eval_t.push_back(j);
    for (int i = 0; i < Kt; ++i) {
      std::memset(&e, 0, sizeof(e));
      e.mode = NODE_ULA; e.gscale = s->ula_grad_scale[j]; e.ss = s->step_sizes[j]; e.std = sqrtf(2.0f * e.ss);
      e.pin = i == Kt - 1; e.hist_slot = i == Kt - 1 ? hslot : -1; e.draw = draw++;
      sched.push_back(e); eval_t.push_back(j);
    }
  }
  const int num_evals = (int)eval_t.size();
  if (p->persist_timed_evals > 0) { int r = drain_timing(p); if (r) return r; }
  // ---- device resources (kept on the plan) ---------------------------------------------------------------------------------
  if (!p->ps_node) {
    CCSP_CUDA_TRY(cudaStreamCreateWithFlags(&p->ps_node, cudaStreamNonBlocking));
    CCSP_CUDA_TRY(cudaStreamCreateWithFlags(&p->ps_edge, cudaStreamNonBlocking));
    CCSP_CUDA_TRY(cudaEventCreateWithFlags(&p->pe_begin, cudaEventDisableTiming));
    CCSP_CUDA_TRY(cudaEventCreateWithFlags(&p->pe_node, cudaEventDisableTiming));
    CCSP_CUDA_TRY(cudaEventCreateWithFlags(&p->pe_edge, cudaEventDisableTiming));
    CCSP_CUDA_TRY(cudaEventCreate(&p->pe_t0));
    CCSP_CUDA_TRY(cudaEventCreate(&p->pe_t1));
    CCSP_CUDA_TRY(p->pool.alloc(&p->p_flags, 64 * CCSP_MAX_CHAINS));
  }
  if (p->p_sched_cap < sched.size()) {
    CCSP_CUDA_TRY(cudaStreamSynchronize(user_st));
    if (p->p_sched) { p->pool.release(p->p_sched); p->pool.release(p->p_eval_t); }
    CCSP_CUDA_TRY(p->pool.alloc(&p->p_sched, sched.size()));
    CCSP_CUDA_TRY(p->pool.alloc(&p->p_eval_t, sched.size()));
    p->p_sched_cap = sched.size();
  }
  CCSP_CUDA_TRY(cudaMemcpyAsync(p->p_sched, sched.data(), sched.size() * sizeof(NodeEval), cudaMemcpyHostToDevice, user_st));
  CCSP_CUDA_TRY(cudaMemcpyAsync(p->p_eval_t, eval_t.data(), eval_t.size() * sizeof(int), cudaMemcpyHostToDevice, user_st));
  CCSP_CUDA_TRY(cudaMemsetAsync(p->p_flags, 0, 64 * CCSP_MAX_CHAINS * sizeof(unsigned), user_st));
  CCSP_CUDA_TRY(cudaEventRecord(p->pe_begin, user_st));
  CCSP_CUDA_TRY(cudaStreamWaitEvent(p->ps_edge, p->pe_begin, 0));
  CCSP_CUDA_TRY(cudaStreamWaitEvent(p->ps_node, p->pe_begin, 0));
  ensure_trap_word();
  const unsigned node_ctas = (unsigned)cfg.node_ctas;
  const int NC = p->num_chains;
  // flag words of chain c: node_done + 32 c, edge_done + 32 c (every word on its own 128-byte line)
  unsigned *node_done = p->p_flags, *edge_done = p->p_flags + 32 * CCSP_MAX_CHAINS;
  // the edge clusters must all be resident before the node CTAs take SMs: they need whole TPCs, and a node CTA that got there
  // first would leave a cluster waiting for a TPC that never frees up (both kernels run until the sample is done)
  ensure_arrival_word(p);
  if (p->arrive_host) *p->arrive_host = 0;
  // both functions are loaded before either runs (lazy module loading would otherwise stall the second launch behind the first kernel)
  CCSP_CUDA_TRY((tc::configure_node_tc_persistent<M>()));
  // ---- the edge kernel first: its clusters take whole TPCs -------------------------------------------------------------------
  {
    tc::FusedArgs a;
    std::memset(&a, 0, sizeof(a));
    a.pe_split = p->pe_split[1];
    a.src_i = p->src_i; a.src_j = p->src_j;
    a.b_blob = m->blob_l1[m->math]; a.w_blob = m->blob_dec[m->math];
    a.tile_type = p->tile_type;
    a.num_m_tiles = (int)(p->Epad / CCSP_TILE_M);
    a.S = p->S; a.tb = m->tb;
    a.bd1 = m->dec_b1; a.Wd2 = m->dec_w2; a.bd2 = m->dec_b2; a.P = m->P; a.o = p->o;
  a.pad_row_plus1 = (int)p->n + 1;
    a.num_evals = num_evals; a.eval_t = p->p_eval_t; a.tb_base = m->tb; a.tb_stride = m->C * CCSP_H2;
    a.node_done = node_done; a.edge_done = edge_done; a.node_ctas = node_ctas;
    a.num_chains = NC;
    for (int c = 0; c <= NC; ++c) a.chain_tile0[c] = p->chain_tile0[c];
    a.arrive = p->arrive_dev;
    for (int c = 0; c < NC; ++c) a.chain_nblk[c] = (unsigned)((p->chain_row0[c + 1] - p->chain_row0[c] + 63) / 64);
    // carry epilogue-2 across chain evaluations only where a pair has several units per chain evaluation to hide it behind
    a.drain_each_eval = (NC == 1 || cfg.partition || getenv("CCSP_PIPE_DRAIN")) ? 1 : 0;
    a.trace = getenv("CCSP_PERSIST_TRACE") ? (long long *)g_ptrace_dev : nullptr;
    if (p->timing_stride > 0) CCSP_CUDA_TRY(cudaEventRecord(p->pe_t0, p->ps_edge));
    CCSP_CUDA_TRY((tc::launch_fused2_persistent<M>(a, pairs, p->ps_edge)));
    if (p->timing_stride > 0) { CCSP_CUDA_TRY(cudaEventRecord(p->pe_t1, p->ps_edge)); p->persist_timed_evals = num_evals; }
    count_launch();
  }
  if (NC > 1 && p->arrive_host) {
    // wait (host) until every edge CTA has started; bounded: after 20 s go on and let the flag waits' own trap decide
    const auto t0 = std::chrono::steady_clock::now();
    while (*(volatile unsigned *)p->arrive_host < (unsigned)(2 * pairs)) {
      if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 20.0) break;
    }
  }
  {
    NodeArgs a = node_args_base(p);
    a.z = nz->noise; a.seed = nz->seed; a.node_offset = nz->node_offset;
    a.has_xinit = nz->x_init != nullptr; a.x_in = nz->x_init;
    a.hist = history;
    a.sched = p->p_sched; a.num_iters = (int)sched.size(); a.nP = nP;
    a.node_done = node_done; a.edge_done = edge_done; a.edge_ctas = (unsigned)(2 * pairs);
    a.num_chains = NC;
    for (int c = 0; c <= NC; ++c) a.chain_row0[c] = p->chain_row0[c];
    for (int c = 0; c < NC; ++c) a.chain_units[c] = (unsigned)(p->chain_tile0[c + 1] - p->chain_tile0[c]);
    a.node_partition = cfg.partition;
    a.trace = getenv("CCSP_PERSIST_TRACE") ? (long long *)g_ptrace_dev : nullptr;
    CCSP_CUDA_TRY((tc::launch_node_tc_persistent<M>(a, m->blob_pose[m->math], p->ps_node, (int)node_ctas)));
    count_launch();
  }
  CCSP_CUDA_TRY(cudaEventRecord(p->pe_edge, p->ps_edge));
  CCSP_CUDA_TRY(cudaEventRecord(p->pe_node, p->ps_node));
  CCSP_CUDA_TRY(cudaStreamWaitEvent(user_st, p->pe_edge, 0));
  CCSP_CUDA_TRY(cudaStreamWaitEvent(user_st, p->pe_node, 0));
  CCSP_CUDA_TRY(cudaMemcpyAsync(out, p->x, nP * sizeof(float), cudaMemcpyDeviceToDevice, user_st));
  return CCSP_OK;
}

// =================================================================================================
// extern "C"
// =================================================================================================
extern "C" {

const char *ccsp_last_error(void) { return g_last_error.c_str(); }
/* developer aid: (line or code << 40 | block << 24 | thread) written by the device thread that trapped, 0 if none */
unsigned long long ccsp_debug_trap_info(void) { return g_trap_host ? g_trap_host[0] : 0ull; }
/* developer aid (CCSP_PERSIST_TRACE=1): timeline word [event 0..7][iteration 0..31] of the last persistent sample, ns */
unsigned long long ccsp_debug_persist_trace(int event, int iter) {
  return (g_trap_host && event >= 0 && event < 12 && iter >= 0 && iter < 32) ? g_trap_host[8 + event * 32 + iter] : 0ull;
}
/* developer aid / host-logic test hook (no device needed): the scene-group boundaries ccsp_plan_create would use for `want` chains */
int ccsp_debug_chain_cuts(const int64_t *edge_index, const float *edge_attr, int64_t n, int64_t E, int32_t num_types, int32_t want,
                          int64_t *bounds_out) {
  if (!bounds_out || n <= 0 || E < 0 || (E > 0 && (!edge_index || !edge_attr)) || want < 1 || want > CCSP_MAX_CHAINS) return -1;
  std::vector<int> etype(E);
  for (int64_t e = 0; e < E; ++e) {
    const float a = edge_attr[e];
    int c = (a >= 0.f && a < (float)num_types) ? (int)a : -1;
    if (c >= 0 && (float)c != a) c = -1;
    if (c >= 0 && (edge_index[e] < 0 || edge_index[e] >= n || edge_index[E + e] < 0 || edge_index[E + e] >= n)) return -1;
    etype[e] = c;
  }
  const std::vector<int64_t> b = chain_cuts(n, E, edge_index, etype, want);
  for (size_t i = 0; i < b.size(); ++i) bounds_out[i] = b[i];
  return (int)b.size() - 1;
}
int ccsp_abi_version(void) { return CCSP_ABI_VERSION; }
uint64_t ccsp_launch_count(void) { return g_launches; }
void ccsp_reset_launch_count(void) { g_launches = 0; }

int ccsp_model_create(const CcspModelDesc *d, CcspModel **out) {
  CCSP_REQUIRE(d && out, "null descriptor/output");
  CCSP_REQUIRE(d->hidden_dim == CCSP_HIDDEN_DIM, "hidden_dim must be 256");
  CCSP_REQUIRE(d->pose_dim >= 1 && d->pose_dim <= CCSP_MAX_POSE_DIM, "pose_dim out of range");
  CCSP_REQUIRE(d->geom_dim >= 1 && d->geom_dim <= CCSP_MAXP, "geom_dim out of range (1..8)");
  CCSP_REQUIRE(d->grasp_dim >= 0 && d->grasp_dim <= CCSP_MAXP, "grasp_dim out of range (0..8)");
  CCSP_REQUIRE(d->num_types >= 1 && d->num_types <= CCSP_MAX_TYPES, "num_types out of range");
  CCSP_REQUIRE(d->geom_w0 && d->pose_w0 && d->dec_w0 && d->time_w1 && d->mlp_w && d->mlp_b, "null weight pointer");
  CCSP_REQUIRE(d->grasp_dim == 0 || d->grasp_w0, "grasp encoder weights missing");
  int ndev = 0;
  CCSP_CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (ndev == 0) { set_error("no CUDA device: libccsp_b200 has no CPU fallback"); return CCSP_ERR_CUDA; }
  CcspModel *m = new CcspModel();
  int rc = model_build(m, d);
  if (rc != CCSP_OK) { m->pool.free_all(); delete m; return rc; }
  *out = m;
  return CCSP_OK;
}

void ccsp_model_destroy(CcspModel *m) {
  if (!m) return;
  for (CcspPlan *p : m->plans) { p->pool.cache = nullptr; p->m = nullptr; }   // plans destroyed later free their own blocks
  m->cache.clear();
  m->pool.free_all();
  delete m;
}

int ccsp_model_set_math(CcspModel *m, int math) {
  CCSP_REQUIRE(m, "null model");
  CCSP_REQUIRE(math >= CCSP_MATH_FP32 && math <= CCSP_MATH_BF16, "unknown math mode");
  CCSP_CUDA_TRY(cudaSetDevice(m->device));
  int rc = ensure_tc_blobs(m, math);
  if (rc) return rc;
  m->math = math;
  return CCSP_OK;
}

int ccsp_model_get_math(const CcspModel *m) { return m ? m->math : CCSP_ERR_INVALID; }

int ccsp_plan_create(CcspModel *m, const float *x, int64_t n, int32_t F, const int64_t *edge_index,
                     const float *edge_attr, const int8_t *mask, int64_t E, int32_t pose_begin,
                     int32_t grasp_begin, void *stream, CcspPlan **out) {
  CCSP_REQUIRE(m && x && mask && out, "null argument");
  CCSP_REQUIRE(n > 0 && n < (1ll << 23) && E >= 0 && E < (1ll << 22), "n/E out of range (32-bit row offsets)");
  CCSP_REQUIRE(E == 0 || (edge_index && edge_attr), "null edge arrays");
  const int P = m->P, G = m->G, C = m->C;
  CCSP_REQUIRE(F >= G && pose_begin >= 0 && pose_begin + P <= F, "feature slices exceed row width");
  CCSP_REQUIRE(m->Gr == 0 || (grasp_begin >= 0 && grasp_begin + m->Gr <= F), "grasp slice exceeds row width");
  cudaStream_t st = (cudaStream_t)stream;
  CCSP_CUDA_TRY(cudaSetDevice(m->device));

  // ---- host: group edges by type (stable), pad each type to whole 128-row tiles ----------------
  // denoise_fn.py:317: `edge_attr == i` on a float tensor; ids outside [0,C) never match any type.
  std::vector<int> etype(E);
  std::vector<int64_t> cnt(C, 0);
  for (int64_t e = 0; e < E; ++e) {
    float a = edge_attr[e];
    int c = (a >= 0.f && a < (float)C) ? (int)a : -1;
    if (c >= 0 && (float)c != a) c = -1;
    etype[e] = c;
    if (c >= 0) {
      int64_t i = edge_index[e], j = edge_index[E + e];
      CCSP_REQUIRE(i >= 0 && i < n && j >= 0 && j < n, "edge_index out of range");
      ++cnt[c];
    }
  }
  // ---- chains: independent groups of whole scenes (contiguous node ranges that no edge crosses) --------------------------
  // The pipelined sampling path (sample_persistent) interleaves the chains: while the node kernel updates the nodes of one
  // chain, the edge kernel works on another.  Rows are grouped (chain, type); within a node the accumulation order is
  // unchanged (a node's edges all lie in its own chain, still type-major in edge order).
  int want_chains = 1;
  // Opt-in (CCSP_CHAINS=2..4): measured on B200 the pipelined path is no faster than two launches per evaluation at any
  // batch size (profiles/README.md R2.8) — the SMs it takes from the edge phase for the node CTAs cost what the hidden
  // node phase and ramps save — and every extra chain pads every type once more (+2 % rows per chain at config 2).
  if (const char *env = getenv("CCSP_CHAINS")) want_chains = std::max(1, std::min(CCSP_MAX_CHAINS, atoi(env)));
  const std::vector<int64_t> chain_node0 = chain_cuts(n, E, edge_index, etype, want_chains);
  const int NC = (int)chain_node0.size() - 1;
  auto chain_of = [&](int64_t v) { int c = 0; while (c + 1 < NC && v >= chain_node0[c + 1]) ++c; return c; };
  std::vector<int64_t> gcnt((size_t)NC * C, 0);
  for (int64_t e = 0; e < E; ++e)
    if (etype[e] >= 0) ++gcnt[(size_t)chain_of(edge_index[e]) * C + etype[e]];
  std::vector<int64_t> start((size_t)NC * C + 1, 0);
  std::vector<int> tile_type;
  std::vector<int> chain_tile0(NC + 1, 0);
  for (int g = 0; g < NC * C; ++g) {
    int64_t tiles = (gcnt[g] + CCSP_PAD_M - 1) / CCSP_PAD_M * CCSP_CLUSTER;   // whole cluster groups of 128-row tiles
    start[g + 1] = start[g] + tiles * CCSP_TILE_M;
    for (int64_t k = 0; k < tiles; ++k) tile_type.push_back(g % C);
    if (g % C == C - 1) chain_tile0[g / C + 1] = (int)tile_type.size();
  }
  const int64_t Epad = start[(size_t)NC * C];
  std::vector<int> src_i(Epad, (int)n), src_j(Epad, (int)n);   // padded rows read the zero row n
  {
    std::vector<int64_t> fill(start.begin(), start.end() - 1);
    for (int64_t e = 0; e < E; ++e) {
      int c = etype[e];
      if (c < 0) continue;
      int64_t pos = fill[(size_t)chain_of(edge_index[e]) * C + c]++;
      src_i[pos] = (int)edge_index[e];
      src_j[pos] = (int)edge_index[E + e];
    }
  }
  // destination CSR in the reference's accumulation order (type-major, edge order, arg1 before arg2;
  // denoise_fn.py:380-383 scatter_add_ over args.reshape(-1), types visited in order at :512)
  std::vector<int> node_ptr(n + 1, 0);
  for (int64_t pos = 0; pos < Epad; ++pos) {
    if (src_i[pos] < n) { ++node_ptr[src_i[pos] + 1]; ++node_ptr[src_j[pos] + 1]; }
  }
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
  std::vector<float> gt((size_t)n * P), xtail((size_t)n * P);
  for (int64_t v = 0; v < n; ++v)
    for (int p = 0; p < P; ++p) {
      gt[v * P + p] = x[v * F + pose_begin + p];      // ddpm.py:269
      xtail[v * P + p] = x[v * F + (F - P) + p];      // denoise_fn.py:533 (last P columns)
    }
  std::vector<signed char> hmask(mask, mask + n);
  for (auto &b : hmask) b = b != 0;

  // ---- device: upload + static term --------------------------------------------------------------
  g_upload_bytes = 0;
  CcspPlan *p = new CcspPlan();
  p->pool.cache = &m->cache;
  p->m = m; p->device = m->device; p->n = n; p->E = E; p->Epad = Epad; p->num_tiles = (int)tile_type.size();
  p->num_chains = NC;
  for (int c = 0; c <= NC; ++c) { p->chain_tile0[c] = chain_tile0[c]; p->chain_row0[c] = (int)chain_node0[c]; }
  auto fail = [&](int rc) { p->pool.free_all(); delete p; return rc; };
#define PLAN_TRY(expr)                                                                             \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e));                        \
      return fail(CCSP_ERR_CUDA);                                                                  \
    }                                                                                              \
  } while (0)
  PLAN_TRY(upload(p->pool, &p->src_i, src_i));
  PLAN_TRY(upload(p->pool, &p->src_j, src_j));
  PLAN_TRY(upload(p->pool, &p->tile_type, tile_type));
  PLAN_TRY(upload(p->pool, &p->node_ptr, node_ptr));
  PLAN_TRY(upload(p->pool, &p->node_src, node_src));
  PLAN_TRY(upload(p->pool, &p->mask, hmask));
  PLAN_TRY(upload(p->pool, &p->gt, gt));
  PLAN_TRY(upload(p->pool, &p->xtail, xtail));
  PLAN_TRY(p->pool.alloc(&p->S, (size_t)Epad * CCSP_H2));
  PLAN_TRY(p->pool.alloc(&p->o, (size_t)Epad * 2 * P));
  PLAN_TRY(p->pool.alloc(&p->x, (size_t)n * P));
  PLAN_TRY(cudaMemsetAsync(p->o, 0, (size_t)Epad * 2 * P * sizeof(float), st));

  if (Epad > 0) {
    float *xdev = nullptr, *ge = nullptr, *gr = nullptr;
    PLAN_TRY(p->pool.alloc(&xdev, (size_t)n * F));
    PLAN_TRY(p->pool.alloc(&ge, (size_t)(n + 1) * CCSP_H));
    PLAN_TRY(cudaMemcpyAsync(xdev, x, (size_t)n * F * sizeof(float), cudaMemcpyHostToDevice, st));
    g_upload_bytes += (int64_t)((size_t)n * F * sizeof(float));
    const unsigned eb = (unsigned)((n + 1 + ENC_ROWS - 1) / ENC_ROWS);
    k_encode_rows<<<eb, 256, 0, st>>>(xdev, F, 0, G, (int)n, (int)n + 1, m->geom.w0, m->geom.b0, m->geom.w2t,
                                      m->geom.b2, ge);
    count_launch();
    RowSrc rs;
    rs.src[2] = nullptr; rs.idx[2] = nullptr;
    if (m->Gr > 0) {
      PLAN_TRY(p->pool.alloc(&gr, (size_t)(n + 1) * CCSP_H));
      k_encode_rows<<<eb, 256, 0, st>>>(xdev, F, grasp_begin, m->Gr, (int)n, (int)n + 1, m->grasp.w0, m->grasp.b0,
                                        m->grasp.w2t, m->grasp.b2, gr);
      count_launch();
      rs.nseg = 3;
      rs.src[0] = gr; rs.idx[0] = p->src_i;      // grasp_emb[args_1]   denoise_fn.py:337
      rs.src[1] = ge; rs.idx[1] = p->src_i;
      rs.src[2] = ge; rs.idx[2] = p->src_j;
    } else {
      rs.nseg = 2;
      rs.src[0] = ge; rs.idx[0] = p->src_i;
      rs.src[1] = ge; rs.idx[1] = p->src_j;
    }
    k_edge_l1_simt<EPI_STATIC><<<dim3((unsigned)(Epad / SG_BM), CCSP_H2 / SG_BN), 256, 0, st>>>(
        rs, m->Wst, p->tile_type, m->bias, nullptr, nullptr, p->S);
    count_launch();
    PLAN_TRY(cudaGetLastError());
    PLAN_TRY(cudaStreamSynchronize(st));
    p->pool.release(xdev);
    p->pool.release(ge);
    if (gr) p->pool.release(gr);
  } else {
    PLAN_TRY(cudaStreamSynchronize(st));
  }
#undef PLAN_TRY
  p->h2d_bytes = g_upload_bytes;
  m->plans.push_back(p);
  *out = p;
  return CCSP_OK;
}

void ccsp_plan_destroy(CcspPlan *p) {
  if (!p) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(p->device);       // the caller's current device need not be the plan's
  cudaDeviceSynchronize();        // the blocks go back to the model's cache and may be reused at once
  if (p->m) {
    auto &v = p->m->plans;
    for (size_t i = 0; i < v.size(); ++i)
      if (v[i] == p) { v.erase(v.begin() + i); break; }
  }
  for (cudaEvent_t e : p->ev) cudaEventDestroy(e);
  if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
  if (p->capture_stream) cudaStreamDestroy(p->capture_stream);
  if (p->ps_node) cudaStreamDestroy(p->ps_node);
  if (p->ps_edge) cudaStreamDestroy(p->ps_edge);
  if (p->pe_begin) cudaEventDestroy(p->pe_begin);
  if (p->pe_node) cudaEventDestroy(p->pe_node);
  if (p->pe_edge) cudaEventDestroy(p->pe_edge);
  if (p->pe_t0) cudaEventDestroy(p->pe_t0);
  if (p->pe_t1) cudaEventDestroy(p->pe_t1);
  if (p->arrive_host) cudaFreeHost(p->arrive_host);
  p->pool.free_all();
  delete p;
  cudaSetDevice(prev);
}

int64_t ccsp_plan_num_nodes(const CcspPlan *p) { return p ? p->n : -1; }
int64_t ccsp_plan_num_edges(const CcspPlan *p) { return p ? p->E : -1; }
int64_t ccsp_plan_num_edge_rows(const CcspPlan *p) { return p ? p->Epad : -1; }
int64_t ccsp_plan_h2d_bytes(const CcspPlan *p) { return p ? p->h2d_bytes : -1; }

int ccsp_plan_set_timing(CcspPlan *p, int32_t stride) {
  CCSP_REQUIRE(p && stride >= 0, "bad timing stride");
  p->timing_stride = stride;
  const size_t cap = 4 * 1024;
  if (stride > 0 && p->ev.empty()) {
    p->ev.resize(cap);
    for (auto &e : p->ev) CCSP_CUDA_TRY(cudaEventCreate(&e));
  }
  return CCSP_OK;
}

static int drain_timing(CcspPlan *p) {
  for (size_t g = 0; g + 4 <= p->ev_used; g += 4) {
    float a = 0, b = 0, c = 0;
    CCSP_CUDA_TRY(cudaEventSynchronize(p->ev[g + 3]));
    if (edge_fused(p->m->math)) {          // one edge kernel: everything between the first and third event
      CCSP_CUDA_TRY(cudaEventElapsedTime(&a, p->ev[g], p->ev[g + 2]));
    } else {
      CCSP_CUDA_TRY(cudaEventElapsedTime(&a, p->ev[g], p->ev[g + 1]));
      CCSP_CUDA_TRY(cudaEventElapsedTime(&b, p->ev[g + 1], p->ev[g + 2]));
    }
    CCSP_CUDA_TRY(cudaEventElapsedTime(&c, p->ev[g + 2], p->ev[g + 3]));
    p->timing.samples += 1; p->timing.ms_edge_l1 += a; p->timing.ms_edge_dec += b; p->timing.ms_node += c;
  }
  p->ev_used = 0;
  if (p->persist_timed_evals > 0) {        // persistent edge kernel: its whole duration, booked as that many evaluations
    float a = 0;
    CCSP_CUDA_TRY(cudaEventSynchronize(p->pe_t1));
    CCSP_CUDA_TRY(cudaEventElapsedTime(&a, p->pe_t0, p->pe_t1));
    p->timing.samples += p->persist_timed_evals; p->timing.ms_edge_l1 += a;
    p->persist_timed_evals = 0;
  }
  return CCSP_OK;
}

int ccsp_plan_get_timing(CcspPlan *p, CcspTiming *out) {
  CCSP_REQUIRE(p && out, "null argument");
  int rc = drain_timing(p);
  if (rc) return rc;
  *out = p->timing;
  p->timing = CcspTiming{0, 0.0, 0.0, 0.0};
  return CCSP_OK;
}

int ccsp_denoise(CcspPlan *p, const float *poses, int32_t t, float *out, void *stream) {
  CCSP_REQUIRE(p && poses && out, "null argument");
  CCSP_REQUIRE(t >= 0 && t < (1 << 20), "timestep out of range");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if ((rc = ensure_time_table(p->m, t + 1, st))) return rc;
  if ((rc = ensure_mode_buffers(p))) return rc;
  NodeArgs a = node_args_base(p);
  a.mode = NODE_ENCODE; a.x_in = poses;
  if ((rc = launch_node(p, a, st))) return rc;
  if ((rc = launch_edge(p, t, st))) return rc;
  a = node_args_base(p);
  a.mode = NODE_EPS_OUT; a.eps_out = out;
  return launch_node(p, a, st);
}

int ccsp_sample(CcspPlan *p, const CcspSchedule *s, const CcspNoise *nz, float *out, float *history,
                void *stream) {
  CCSP_REQUIRE(p && s && nz && out, "null argument");
  CCSP_REQUIRE(s->T >= 1, "T must be >= 1");
  CCSP_REQUIRE(s->sqrt_recip_alphas_cumprod && s->sqrt_recipm1_alphas_cumprod && s->posterior_mean_coef1 &&
                   s->posterior_mean_coef2 && s->posterior_log_variance_clipped, "null schedule table");
  CCSP_REQUIRE(!s->samples_per_step || (s->ula_grad_scale && s->step_sizes), "ULA tables missing");
  cudaStream_t user_st = (cudaStream_t)stream, st = user_st;
  CcspModel *m = p->m;
  const int T = s->T;
  const int per = s->ebm_per_steps > 0 ? s->ebm_per_steps : 1;
  const size_t nP = (size_t)p->n * m->P;
  int rc;
  if ((rc = ensure_time_table(m, T, st))) return rc;
  if ((rc = ensure_mode_buffers(p))) return rc;
  {
    PersistCfg cfg;
    if (persistent_eligible(p, &cfg)) return sample_persistent(p, s, nz, out, history, user_st, cfg);
  }

  // sampled kernel timing: evaluation `ev_idx` is bracketed by events when selected
  long ev_idx = 0;
  auto timed_eval = [&](int t, NodeArgs &na) -> int {
    const bool sel = p->timing_stride > 0 && (ev_idx++ % p->timing_stride) == 0;
    if (sel && p->ev_used + 4 > p->ev.size()) { int r = drain_timing(p); if (r) return r; }
    cudaEvent_t *e = sel ? &p->ev[p->ev_used] : nullptr;
    int r;
    if (e) CCSP_CUDA_TRY(cudaEventRecord(e[0], st));
    if ((r = launch_edge(p, t, st, e ? e[1] : nullptr))) return r;
    if (e) CCSP_CUDA_TRY(cudaEventRecord(e[2], st));
    if ((r = launch_node(p, na, st))) return r;
    if (e) { CCSP_CUDA_TRY(cudaEventRecord(e[3], st)); p->ev_used += 4; }
    return CCSP_OK;
  };

  unsigned draw = 0;
  auto zptr = [&](unsigned d) { return nz->noise ? nz->noise + (size_t)d * nP : (const float *)nullptr; };
  auto with_noise = [&](NodeArgs &a) {
    a.z = zptr(draw); a.seed = nz->seed; a.node_offset = nz->node_offset; a.draw = draw;
    ++draw;
  };

  // Optional (CCSP_GRAPH=1): capture the whole launch sequence into one CUDA graph and replay it.  Measured on B200
  // (profiles/README.md, round 2): the loop is bound by the kernels' own latency, not by launch overhead.
  static const bool want_graph = getenv("CCSP_GRAPH") != nullptr;
  const bool use_graph = want_graph && p->timing_stride == 0;
  if (use_graph) {
    if (p->graph_exec) { CCSP_CUDA_TRY(cudaStreamSynchronize(user_st)); cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }
    // the legacy default stream cannot be captured: record on a private stream, replay on the caller's
    if (!p->capture_stream) CCSP_CUDA_TRY(cudaStreamCreateWithFlags(&p->capture_stream, cudaStreamNonBlocking));
    st = p->capture_stream;
    CCSP_CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  }

  // x_T = 0.5 * randn, pinned rows <- gt                                   (ddpm.py:273-274)
  NodeArgs a = node_args_base(p);
  a.mode = NODE_INIT; a.pin = 1; a.has_xinit = nz->x_init != nullptr; a.x_in = nz->x_init;
  a.hist = history;
  with_noise(a);                                   // the draw is consumed even when x_init is given
  if ((rc = launch_node(p, a, st))) return rc;

  for (int j = T - 1; j >= 0; --j) {               // ddpm.py:325
    const int Kt = (s->samples_per_step && (j % per == 0)) ? s->samples_per_step[j] : 0;
    float *hslot = history ? history + (size_t)(T - j) * nP : nullptr;
    // p_sample (ddpm.py:253-258)
    a = node_args_base(p);
    a.mode = NODE_DDPM;
    a.a = s->sqrt_recip_alphas_cumprod[j];
    a.b = s->sqrt_recipm1_alphas_cumprod[j];
    a.c1 = s->posterior_mean_coef1[j];
    a.c2 = s->posterior_mean_coef2[j];
    a.sigma = (j == 0 ? 0.f : 1.f) * expf(0.5f * s->posterior_log_variance_clipped[j]);
    a.pin = Kt == 0;
    a.hist = Kt == 0 ? hslot : nullptr;
    with_noise(a);
    if ((rc = timed_eval(j, a))) return rc;
    // AnnealedULASampler.sample_step (ddpm.py:955-966); rows are re-pinned only after the K steps (:334)
    for (int i = 0; i < Kt; ++i) {
      a = node_args_base(p);
      a.mode = NODE_ULA;
      a.gscale = s->ula_grad_scale[j];
      a.ss = s->step_sizes[j];
      a.std = sqrtf(2.0f * a.ss);
      a.pin = i == Kt - 1;
      a.hist = i == Kt - 1 ? hslot : nullptr;
      with_noise(a);
      if ((rc = timed_eval(j, a))) return rc;
    }
  }
  CCSP_CUDA_TRY(cudaMemcpyAsync(out, p->x, nP * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (use_graph) {
    cudaGraph_t graph = nullptr;
    CCSP_CUDA_TRY(cudaStreamEndCapture(st, &graph));
    CCSP_CUDA_TRY(cudaGraphInstantiate(&p->graph_exec, graph, 0));
    cudaGraphDestroy(graph);
    CCSP_CUDA_TRY(cudaGraphLaunch(p->graph_exec, user_st));
  }
  return CCSP_OK;
}

int ccsp_check_solved(const CcspCheckDesc *d, const float *poses, uint8_t *solved, int32_t *counts, void *stream) {
  CCSP_REQUIRE(d && poses && solved, "null argument");
  CCSP_REQUIRE(d->kind == CCSP_WORLD_BOXES || d->kind == CCSP_WORLD_QUALITATIVE, "unknown world kind");
  CCSP_REQUIRE(d->num_scenes >= 0 && d->x && d->scene_node_ptr && d->world_dims, "null scene arrays");
  CCSP_REQUIRE(d->kind == CCSP_WORLD_BOXES || (d->scene_edge_ptr && d->edge_a && d->edge_b && d->edge_type), "null edge arrays");
  CCSP_REQUIRE(d->P >= 2 && d->pose_begin >= 2 && d->pose_begin + d->P <= d->F, "rows must be [w, l, ..., pose...]");
  CCSP_REQUIRE(d->kind == CCSP_WORLD_BOXES ? (d->F == 4 && d->pose_begin == 2 && d->P == 2)
                                           : (d->F == 6 && d->pose_begin == 2 && d->P == 4),
               "row layout: boxes [w,l,x,y]; qualitative [w,l,x,y,cs,sn] (data_utils.py:229-258)");
  if (d->num_scenes == 0) return CCSP_OK;
  int ndev = 0;
  CCSP_CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (ndev == 0) { set_error("no CUDA device: libccsp_b200 has no CPU fallback"); return CCSP_ERR_CUDA; }
  CheckArgs a;
  a.kind = d->kind; a.S = d->num_scenes; a.F = d->F; a.P = d->P; a.pose_begin = d->pose_begin; a.clamp = d->clamp;
  a.x = d->x; a.poses = poses; a.world_dims = d->world_dims;
  a.scene_node_ptr = d->scene_node_ptr; a.scene_edge_ptr = d->scene_edge_ptr;
  a.edge_a = d->edge_a; a.edge_b = d->edge_b; a.edge_type = d->edge_type;
  a.solved = solved; a.counts = counts;
  k_check_solved<<<(unsigned)((d->num_scenes + CHECK_WARPS - 1) / CHECK_WARPS), CHECK_WARPS * 32, 0, (cudaStream_t)stream>>>(a);
  CCSP_LAUNCH_CHECK();
  return CCSP_OK;
}

}  // extern "C"
