// tc_gemm_test.cu — standalone numerics + throughput check of the tcgen05 edge kernels (kernels_tc.cuh)
// against a double-precision CPU reference.  Built by `python -m diffusion_ccsp_b200.build --tests`, run on
// the B200 box:   timeout 200 diffusion_ccsp_b200/lib/tc_gemm_test [perf]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../kernels_tc.cuh"
#include "../kernels_fused2.cuh"
#include "kernels_fused3.cuh"

namespace ccsp {
void set_error(const std::string &) {}
void count_launch() {}
}  // namespace ccsp

using namespace ccsp;
using namespace ccsp::tc;

#define CK(x)                                                                                 \
  do {                                                                                        \
    cudaError_t e_ = (x);                                                                     \
    if (e_ != cudaSuccess) {                                                                  \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(2);                                                                                \
    }                                                                                         \
  } while (0)

template <typename T>
T *dev(const std::vector<T> &h) {
  T *d;
  CK(cudaMalloc(&d, h.size() * sizeof(T) + 16));
  CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

static double silu_d(double x) { return x / (1.0 + exp(-x)); }

struct Problem {
  int n_nodes, m_tiles, groups, P;
  std::vector<float> node_emb;       // [n_nodes+1][256]
  std::vector<int> idx0, idx1, tile_type;
  std::vector<float> W;              // [groups][512][512]
  std::vector<float> S, tb;          // [rows][512] (row-major here; blocked for the device), [groups][512]
  std::vector<float> Wd1, bd1, Wd2, bd2;   // [128][256], [128], [P][128], [P]
};

static Problem make_problem(int n_nodes, int m_tiles, int groups, unsigned seed) {
  Problem p;
  p.n_nodes = n_nodes; p.m_tiles = m_tiles; p.groups = groups; p.P = 4;
  std::mt19937 rng(seed);
  std::uniform_real_distribution<float> u(-1.f, 1.f);
  const int rows = m_tiles * 128;
  p.node_emb.resize((size_t)(n_nodes + 1) * 256);
  for (auto &v : p.node_emb) v = u(rng);
  for (int k = 0; k < 256; ++k) p.node_emb[(size_t)n_nodes * 256 + k] = 0.f;
  p.idx0.resize(rows); p.idx1.resize(rows);
  for (int r = 0; r < rows; ++r) { p.idx0[r] = rng() % (n_nodes + 1); p.idx1[r] = rng() % (n_nodes + 1); }
  p.tile_type.resize(m_tiles);
  for (int t = 0; t < m_tiles; ++t) p.tile_type[t] = (t / 4) % groups;     // constant within cluster groups of up to 4 tiles
  p.W.resize((size_t)groups * 512 * 512);
  for (auto &v : p.W) v = u(rng) * 0.0442f;
  p.S.resize((size_t)rows * 512);
  for (auto &v : p.S) v = u(rng);
  p.tb.resize((size_t)groups * 512);
  for (auto &v : p.tb) v = u(rng);
  p.Wd1.resize(128 * 256); for (auto &v : p.Wd1) v = u(rng) * 0.0625f;
  p.bd1.resize(128); for (auto &v : p.bd1) v = u(rng) * 0.0625f;
  p.Wd2.resize(p.P * 128); for (auto &v : p.Wd2) v = u(rng) * 0.088f;
  p.bd2.resize(p.P); for (auto &v : p.bd2) v = u(rng) * 0.088f;
  return p;
}

// host mirror of the node kernel's split pose-embedding rows
template <class M>
static std::vector<uint8_t> split_rows(const std::vector<float> &emb) {
  const size_t rows = emb.size() / 256;
  std::vector<uint8_t> out(rows * M::PE_ROW_BYTES);
  for (size_t r = 0; r < rows; ++r)
    for (int k = 0; k < 256; ++k) {
      const float x = emb[r * 256 + k];
      uint8_t *row = out.data() + r * M::PE_ROW_BYTES;
      if (M::KIND == KIND_TF32) {
        uint32_t hi = host_tf32_rna(x);
        float hf; memcpy(&hf, &hi, 4);
        uint32_t lo = host_tf32_rna(x - hf);
        memcpy(row + M::pe_elem_off(k, 0), &hi, 4); memcpy(row + M::pe_elem_off(k, 1), &lo, 4);
      } else {
        uint16_t hi = host_bf16_rn(x), lo = host_bf16_rn(x - host_bf16_to_f(hi));
        memcpy(row + M::pe_elem_off(k, 0), &hi, 2); memcpy(row + M::pe_elem_off(k, 1), &lo, 2);
      }
    }
  return out;
}

// unpack operand-format H back to row-major floats [rows][512] (hi + lo)
template <class M>
static std::vector<float> unpack_H(const std::vector<uint8_t> &Hop, int m_tiles) {
  std::vector<float> H((size_t)m_tiles * 128 * 512);
  const int EPC = 16 / M::ELT;
  for (int tile = 0; tile < m_tiles * 2; ++tile)
    for (int kc = 0; kc < M::NKC2; ++kc)
      for (int r = 0; r < 128; ++r)
        for (int kk = 0; kk < M::KC; ++kk) {
          const uint8_t *st = Hop.data() + (size_t)tile * M::H_TILE_BYTES + (size_t)kc * M::A_STAGE;
          const uint32_t off = sw64_off(r, kk / EPC) + (kk % EPC) * M::ELT;
          float v = 0.f;
          for (int part = 0; part < M::NS; ++part) {
            if (M::KIND == KIND_TF32) { float f; memcpy(&f, st + part * PART + off, 4); v += f; }
            else { uint16_t h; memcpy(&h, st + part * PART + off, 2); v += host_bf16_to_f(h); }
          }
          H[((size_t)(tile >> 1) * 128 + r) * 512 + (tile & 1) * 256 + kc * M::KC + kk] = v;
        }
  return H;
}

struct Report { double max_err, max_ref, ms; };

static void compare(const float *got, const std::vector<double> &ref, Report &rep, const char *what) {
  int shown = 0;
  for (size_t i = 0; i < ref.size(); ++i) {
    double e = fabs((double)got[i] - ref[i]);
    if (!(e == e)) e = 1e30;
    if (e > rep.max_err) rep.max_err = e;
    if (fabs(ref[i]) > rep.max_ref) rep.max_ref = fabs(ref[i]);
    if (e > 0.05 && shown < 4) { printf("    %s mismatch at %zu: got %g want %g\n", what, i, got[i], ref[i]); ++shown; }
  }
}

// runs L1 then DEC; returns errors of H (vs ref) in `h` and of o in `o`
template <class M, int CL = 1>
static void run_chain(const Problem &p, const std::vector<double> *Href, const std::vector<double> *oref, int iters, int sms,
                      Report &h, Report &o, int dbg = 0) {
  const int rows = p.m_tiles * 128;
  std::vector<float> Sblk((size_t)rows * 512);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < 512; ++c) Sblk[blk_off(r, c)] = p.S[(size_t)r * 512 + c];
  std::vector<uint8_t> pe = split_rows<M>(p.node_emb);
  const size_t per = (size_t)2 * M::NKC1 * L1Cfg<M>::B_STAGE;
  std::vector<uint8_t> b1(per * p.groups), b2((size_t)M::NKC2 * DecCfg<M>::B_STAGE);
  for (int g = 0; g < p.groups; ++g) pack_b_blob<M, 256>(&p.W[(size_t)g * 512 * 512], 512, 0, 512, 512, b1.data() + g * per);
  pack_b_blob<M, 128>(p.Wd1.data(), 256, 0, 256, 128, b2.data());
  uint8_t *d_pe = dev(pe), *d_b1 = dev(b1), *d_b2 = dev(b2), *d_H;
  float *d_S = dev(Sblk), *d_tb = dev(p.tb), *d_bd1 = dev(p.bd1), *d_w2 = dev(p.Wd2), *d_bd2 = dev(p.bd2), *d_o;
  int *d_i0 = dev(p.idx0), *d_i1 = dev(p.idx1), *d_tt = dev(p.tile_type);
  const size_t Hbytes = (size_t)p.m_tiles * 2 * M::H_TILE_BYTES;
  CK(cudaMalloc(&d_H, Hbytes));
  CK(cudaMemset(d_H, 0xFF, Hbytes));
  CK(cudaMalloc(&d_o, (size_t)rows * 2 * p.P * sizeof(float)));
  CK(cudaMemset(d_o, 0xFF, (size_t)rows * 2 * p.P * sizeof(float)));
  L1Args a;
  memset(&a, 0, sizeof(a));
  a.pe_split = d_pe; a.src_i = d_i0; a.src_j = d_i1; a.b_blob = d_b1; a.tile_type = d_tt; a.num_m_tiles = p.m_tiles;
  a.S = d_S; a.tb = d_tb; a.H = d_H; a.dbg = dbg;
  DecArgs d;
  memset(&d, 0, sizeof(d));
  d.H = d_H; d.b_blob = d_b2; d.num_tiles = p.m_tiles * 2; d.bd1 = d_bd1; d.Wd2 = d_w2; d.bd2 = d_bd2; d.P = p.P; d.o = d_o; d.dbg = dbg;
  CK((launch_l1_tc<M, CL>(a, sms, 0)));
  CK(launch_dec_tc<M>(d, sms, 0));
  CK(cudaDeviceSynchronize());
  h = Report{0, 0, 0}; o = Report{0, 0, 0};
  if (Href) {
    std::vector<uint8_t> Hop(Hbytes);
    CK(cudaMemcpy(Hop.data(), d_H, Hbytes, cudaMemcpyDeviceToHost));
    std::vector<float> H = unpack_H<M>(Hop, p.m_tiles);
    compare(H.data(), *Href, h, "H");
    std::vector<float> oo((size_t)rows * 2 * p.P);
    CK(cudaMemcpy(oo.data(), d_o, oo.size() * sizeof(float), cudaMemcpyDeviceToHost));
    compare(oo.data(), *oref, o, "o");
  }
  if (iters > 0) {
    cudaEvent_t e0, e1, e2;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) CK((launch_l1_tc<M, CL>(a, sms, 0)));
    CK(cudaEventRecord(e1));
    for (int i = 0; i < iters; ++i) CK(launch_dec_tc<M>(d, sms, 0));
    CK(cudaEventRecord(e2));
    CK(cudaEventSynchronize(e2));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1)); h.ms = ms / iters;
    CK(cudaEventElapsedTime(&ms, e1, e2)); o.ms = ms / iters;
  }
  cudaFree(d_pe); cudaFree(d_b1); cudaFree(d_b2); cudaFree(d_H); cudaFree(d_S); cudaFree(d_tb); cudaFree(d_bd1);
  cudaFree(d_w2); cudaFree(d_bd2); cudaFree(d_o); cudaFree(d_i0); cudaFree(d_i1); cudaFree(d_tt);
}

static int g_dbg_or = 0;
static bool g_trace = false;   // OR'ed into the pair kernel's dbg word (256: relay mode)
// fused kernels: o only (PAIR: CTA-pair / cta_group::2 version)
template <class M, int PAIR = 1>   // 1: CTA pair (cp.async + relay), 2: CTA pair (TMA gather4 / tile loads), 3: decoder operand in TMEM
static void run_fused(const Problem &p, const std::vector<double> *oref, int iters, int sms, Report &o, int dbg = 0) {
  const int rows = p.m_tiles * 128;
  std::vector<float> Sblk((size_t)rows * 512);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < 512; ++c) Sblk[blk_off(r, c)] = p.S[(size_t)r * 512 + c];
  std::vector<uint8_t> pe = split_rows<M>(p.node_emb);
  const size_t per = (size_t)2 * M::NKC1 * Fused2Cfg<M>::B1_BLOB_STAGE;
  std::vector<uint8_t> b1(per * p.groups), b2((size_t)M::NKC2 * Fused2Cfg<M>::W_BLOB_STAGE);
  for (int g = 0; g < p.groups; ++g) pack_b_blob<M, 256>(&p.W[(size_t)g * 512 * 512], 512, 0, 512, 512, b1.data() + g * per);
  pack_b_blob<M, 128>(p.Wd1.data(), 256, 0, 256, 128, b2.data());
  uint8_t *d_pe = dev(pe), *d_b1 = dev(b1), *d_b2 = dev(b2);
  float *d_S = dev(Sblk), *d_tb = dev(p.tb), *d_bd1 = dev(p.bd1), *d_w2 = dev(p.Wd2), *d_bd2 = dev(p.bd2), *d_o;
  int *d_i0 = dev(p.idx0), *d_i1 = dev(p.idx1), *d_tt = dev(p.tile_type);
  CK(cudaMalloc(&d_o, (size_t)rows * 2 * p.P * sizeof(float)));
  CK(cudaMemset(d_o, 0xFF, (size_t)rows * 2 * p.P * sizeof(float)));
  FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.pe_split = d_pe; a.src_i = d_i0; a.src_j = d_i1; a.b_blob = d_b1; a.w_blob = d_b2; a.tile_type = d_tt; a.num_m_tiles = p.m_tiles;
  a.S = d_S; a.tb = d_tb; a.bd1 = d_bd1; a.Wd2 = d_w2; a.bd2 = d_bd2; a.P = p.P; a.o = d_o; a.dbg = dbg | (PAIR ? g_dbg_or : 0);
  PairMaps pm;
  memset(&pm, 0, sizeof(pm));
  if (PAIR == 2) CK((make_pair_maps<M>(&pm, d_pe, (int64_t)(pe.size() / M::PE_ROW_BYTES), d_b1, p.groups, d_b2)));
  auto launch = [&]() -> cudaError_t {
    if (PAIR == 3) return launch_fused3_tc<M>(a, sms, 0);
    return launch_fused2_tc<M>(a, sms, 0, PAIR == 2 ? &pm : nullptr);
  };
  long long *d_tr = nullptr;
  if (PAIR && g_trace) {
    CK(cudaMalloc(&d_tr, 8 * 8 * 16 * sizeof(long long)));
    CK(cudaMemset(d_tr, 0, 8 * 8 * 16 * sizeof(long long)));
    a.trace = d_tr;
  }
  CK(launch());
  CK(cudaDeviceSynchronize());
  if (d_tr) {
    std::vector<long long> tr(8 * 8 * 16);
    CK(cudaMemcpy(tr.data(), d_tr, tr.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    const char *names[8] = {"MMA1", "MMA2", "EPI w0", "EPI w12", "M1 wait", "M1 cmt", "G empty", "G issue"};
    long long t0 = tr[0];
    printf("trace (dbg %d), cycles relative to MMA1 unit-0 start; columns = trace slots; kernel entry at %lld\n", a.dbg, tr[15] ? tr[15] - t0 : 0);
    for (int role = 0; role < 8; ++role)
      for (int it = 0; it < (role < 4 ? 8 : 2); ++it) {
        printf("  %-7s u%d:", names[role], it);
        for (int sl = 0; sl < (role < 4 ? 11 : 16); ++sl) { long long v = tr[(role * 8 + it) * 16 + sl]; if (v) printf(" %7lld", v - t0); else printf("       -"); }
        printf("\n");
      }
    a.trace = nullptr;
    cudaFree(d_tr);
  }
  o = Report{0, 0, 0};
  if (oref) {
    std::vector<float> oo((size_t)rows * 2 * p.P);
    CK(cudaMemcpy(oo.data(), d_o, oo.size() * sizeof(float), cudaMemcpyDeviceToHost));
    compare(oo.data(), *oref, o, "o(fused)");
  }
  if (iters > 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) CK(launch());
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1)); o.ms = ms / iters;
  }
  cudaFree(d_pe); cudaFree(d_b1); cudaFree(d_b2); cudaFree(d_S); cudaFree(d_tb); cudaFree(d_bd1);
  cudaFree(d_w2); cudaFree(d_bd2); cudaFree(d_o); cudaFree(d_i0); cudaFree(d_i1); cudaFree(d_tt);
}

int main(int argc, char **argv) {
  const bool pair_only = argc > 1 && !strncmp(argv[1], "pair", 4);
  if (argc > 1 && !strcmp(argv[1], "pairdirect")) g_dbg_or = 256;
  g_trace = argc > 2 && !strcmp(argv[2], "trace");     // only the CTA-pair fused kernel (numerics + perf)
  const bool perf = argc > 1 && (!strcmp(argv[1], "perf") || pair_only);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, smem/block optin %zu\n", prop.name, sms, prop.sharedMemPerBlockOptin);
  int fails = 0;
  {
    // ---- numerics: 20 tiles of 128 edges, 3 weight groups ---------------------------------------------
    Problem p = make_problem(1000, 20, 3, 1);
    const int rows = p.m_tiles * 128;
    std::vector<double> Href((size_t)rows * 512), oref((size_t)rows * 2 * p.P);
    for (int r = 0; r < rows; ++r) {
      const float *a0 = &p.node_emb[(size_t)p.idx0[r] * 256], *a1 = &p.node_emb[(size_t)p.idx1[r] * 256];
      const int g = p.tile_type[r / 128];
      for (int n = 0; n < 512; ++n) {
        const float *w = &p.W[((size_t)g * 512 + n) * 512];
        double acc = 0;
        for (int k = 0; k < 256; ++k) acc += (double)a0[k] * w[k] + (double)a1[k] * w[256 + k];
        Href[(size_t)r * 512 + n] = silu_d(acc + p.S[(size_t)r * 512 + n] + p.tb[g * 512 + n]);
      }
      for (int slot = 0; slot < 2; ++slot) {
        double dd[128];
        for (int j = 0; j < 128; ++j) {
          double acc = 0;
          for (int k = 0; k < 256; ++k) acc += Href[(size_t)r * 512 + slot * 256 + k] * p.Wd1[j * 256 + k];
          dd[j] = silu_d(acc + p.bd1[j]);
        }
        for (int pp = 0; pp < p.P; ++pp) {
          double acc = 0;
          for (int j = 0; j < 128; ++j) acc += dd[j] * p.Wd2[pp * 128 + j];
          oref[((size_t)r * 2 + slot) * p.P + pp] = acc + p.bd2[pp];
        }
      }
    }
    auto chk = [&](const char *name, Report r, double tol) {
      bool ok = r.max_err <= tol * (r.max_ref > 1 ? r.max_ref : 1);
      printf("%-14s max_err %.3e (max|ref| %.3f) tol %.1e  %s\n", name, r.max_err, r.max_ref, tol, ok ? "PASS" : "FAIL");
      if (!ok) ++fails;
    };
    Report h, o;
    if (!pair_only) {
    run_chain<Mode<KIND_TF32, 3>>(p, &Href, &oref, 0, sms, h, o); chk("H tf32x3", h, 3e-6); chk("o tf32x3", o, 3e-6);
    run_chain<Mode<KIND_BF16, 3>>(p, &Href, &oref, 0, sms, h, o); chk("H bf16x3", h, 2e-5); chk("o bf16x3", o, 2e-5);
    run_chain<Mode<KIND_TF32, 1>>(p, &Href, &oref, 0, sms, h, o); chk("H tf32", h, 5e-3); chk("o tf32", o, 5e-3);
    run_chain<Mode<KIND_BF16, 1>>(p, &Href, &oref, 0, sms, h, o); chk("H bf16", h, 4e-2); chk("o bf16", o, 4e-2);
    run_chain<Mode<KIND_BF16, 3>, 2>(p, &Href, &oref, 0, sms, h, o); chk("H bf16x3 cl2", h, 2e-5); chk("o bf16x3 cl2", o, 2e-5);
    run_chain<Mode<KIND_BF16, 3>, 4>(p, &Href, &oref, 0, sms, h, o); chk("H bf16x3 cl4", h, 2e-5); chk("o bf16x3 cl4", o, 2e-5);
    run_chain<Mode<KIND_TF32, 3>, 2>(p, &Href, &oref, 0, sms, h, o); chk("H tf32x3 cl2", h, 3e-6); chk("o tf32x3 cl2", o, 3e-6);
    }
    run_fused<Mode<KIND_BF16, 3>, 1>(p, &oref, 0, sms, o); chk("o pair bf16x3", o, 2e-5);
    run_fused<Mode<KIND_BF16, 1>, 1>(p, &oref, 0, sms, o); chk("o pair bf16", o, 4e-2);
    run_fused<Mode<KIND_BF16, 3>, 3>(p, &oref, 0, sms, o); chk("o pairTMEM bf16x3", o, 2e-5);
    run_fused<Mode<KIND_BF16, 1>, 3>(p, &oref, 0, sms, o); chk("o pairTMEM bf16", o, 4e-2);
    run_fused<Mode<KIND_BF16, 3>, 2>(p, &oref, 0, sms, o); chk("o pairTMA bf16x3", o, 2e-5);
    run_fused<Mode<KIND_BF16, 1>, 2>(p, &oref, 0, sms, o); chk("o pairTMA bf16", o, 4e-2);
  }
  if (perf && fails == 0) {
    // ---- throughput at the config-2 size: 632 edge tiles (80 896 rows), 13 weight groups ---------------
    Problem p = make_problem(9216, 632, 13, 2);
    const double fl1 = 2.0 * 632 * 128 * 512 * 512, fdec = 2.0 * 632 * 256 * 128 * 256;
    auto pr = [&](const char *name, Report h, Report o) {
      printf("%-30s l1 %.3f ms %6.1f TFLOP/s | dec %.3f ms %6.1f TFLOP/s\n", name, h.ms, fl1 / h.ms / 1e9, o.ms, fdec / o.ms / 1e9);
    };
    Report h, o;
    if (pair_only) {
      Report f;
      const double ff = fl1 + fdec;
      run_fused<Mode<KIND_BF16, 3>, 1>(p, nullptr, 20, sms, f); printf("%-30s fused %.3f ms %6.1f TFLOP/s\n", "bf16x3 pair", f.ms, ff / f.ms / 1e9);
      run_fused<Mode<KIND_BF16, 1>, 1>(p, nullptr, 20, sms, f); printf("%-30s fused %.3f ms %6.1f TFLOP/s\n", "bf16 pair", f.ms, ff / f.ms / 1e9);
      for (int dbg : {1, 2, 4, 7, 8, 15}) {
        run_fused<Mode<KIND_BF16, 3>, 1>(p, nullptr, 10, sms, f, dbg);
        printf("bf16x3 pair dbg=%-2d               fused %.3f ms %6.1f TFLOP/s\n", dbg, f.ms, ff / f.ms / 1e9);
      }
      run_fused<Mode<KIND_BF16, 3>, 3>(p, nullptr, 20, sms, f); printf("%-30s fused %.3f ms %6.1f TFLOP/s\n", "bf16x3 pair TMEM-A2", f.ms, ff / f.ms / 1e9);
      run_fused<Mode<KIND_BF16, 1>, 3>(p, nullptr, 20, sms, f); printf("%-30s fused %.3f ms %6.1f TFLOP/s\n", "bf16 pair TMEM-A2", f.ms, ff / f.ms / 1e9);
      for (int dbg : {1, 2, 4, 7, 8, 15}) {
        run_fused<Mode<KIND_BF16, 3>, 3>(p, nullptr, 10, sms, f, dbg);
        printf("bf16x3 pair TMEM-A2 dbg=%-2d       fused %.3f ms %6.1f TFLOP/s\n", dbg, f.ms, ff / f.ms / 1e9);
      }
      run_fused<Mode<KIND_BF16, 3>, 2>(p, nullptr, 20, sms, f); printf("%-30s fused %.3f ms %6.1f TFLOP/s\n", "bf16x3 pair TMA", f.ms, ff / f.ms / 1e9);
      run_fused<Mode<KIND_BF16, 1>, 2>(p, nullptr, 20, sms, f); printf("%-30s fused %.3f ms %6.1f TFLOP/s\n", "bf16 pair TMA", f.ms, ff / f.ms / 1e9);
      for (int dbg : {1, 2, 4, 7, 8, 15}) {
        run_fused<Mode<KIND_BF16, 3>, 2>(p, nullptr, 10, sms, f, dbg);
        printf("bf16x3 pair TMA dbg=%-2d           fused %.3f ms %6.1f TFLOP/s\n", dbg, f.ms, ff / f.ms / 1e9);
      }
      printf(fails ? "RESULT: FAIL (%d)\n" : "RESULT: PASS\n", fails);
      return fails ? 1 : 0;
    }
    run_chain<Mode<KIND_TF32, 3>>(p, nullptr, nullptr, 20, sms, h, o); pr("tf32x3", h, o);
    run_chain<Mode<KIND_BF16, 3>>(p, nullptr, nullptr, 20, sms, h, o); pr("bf16x3", h, o);
    run_chain<Mode<KIND_BF16, 3>, 2>(p, nullptr, nullptr, 20, sms, h, o); pr("bf16x3 cluster 2", h, o);
    run_chain<Mode<KIND_BF16, 3>, 4>(p, nullptr, nullptr, 20, sms, h, o); pr("bf16x3 cluster 4", h, o);
    run_chain<Mode<KIND_TF32, 3>, 2>(p, nullptr, nullptr, 20, sms, h, o); pr("tf32x3 cluster 2", h, o);
    run_chain<Mode<KIND_TF32, 3>, 4>(p, nullptr, nullptr, 20, sms, h, o); pr("tf32x3 cluster 4", h, o);
    run_chain<Mode<KIND_BF16, 1>, 2>(p, nullptr, nullptr, 20, sms, h, o); pr("bf16 cluster 2", h, o);
    run_chain<Mode<KIND_BF16, 1>, 4>(p, nullptr, nullptr, 20, sms, h, o); pr("bf16 cluster 4", h, o);
    run_chain<Mode<KIND_TF32, 1>>(p, nullptr, nullptr, 20, sms, h, o); pr("tf32", h, o);
    run_chain<Mode<KIND_BF16, 1>>(p, nullptr, nullptr, 20, sms, h, o); pr("bf16", h, o);
    {
      Report f;
      const double ff = fl1 + fdec;
      run_fused<Mode<KIND_BF16, 3>, 1>(p, nullptr, 20, sms, f); printf("%-30s fused %.3f ms %6.1f TFLOP/s\n", "bf16x3 pair", f.ms, ff / f.ms / 1e9);
      run_fused<Mode<KIND_BF16, 1>, 1>(p, nullptr, 20, sms, f); printf("%-30s fused %.3f ms %6.1f TFLOP/s\n", "bf16 pair", f.ms, ff / f.ms / 1e9);
      for (int dbg : {1, 2, 4, 7, 8, 15}) {
        run_fused<Mode<KIND_BF16, 3>, 1>(p, nullptr, 10, sms, f, dbg);
        printf("bf16x3 pair dbg=%-2d               fused %.3f ms %6.1f TFLOP/s\n", dbg, f.ms, ff / f.ms / 1e9);
      }
    }
    const char *abl[16] = {"none", "noA", "noB", "noA,noB", "noEpiIO", "noA,noEpiIO", "noB,noEpiIO", "MMA+sync only",
                           "noMMA", "noMMA,noA", "noMMA,noB", "noMMA,noA,noB", "noMMA,noEpiIO", "", "", "sync skeleton"};
    for (int dbg : {1, 2, 4, 7, 8, 15}) {
      char nm[64];
      snprintf(nm, sizeof nm, "tf32x3 [%s]", abl[dbg]);
      run_chain<Mode<KIND_TF32, 3>>(p, nullptr, nullptr, 10, sms, h, o, dbg); pr(nm, h, o);
    }
    for (int dbg : {1, 2, 4, 7, 8, 15, 16, 32, 64, 64 + 3, 16 + 3, 32 + 3}) {
      char nm[64];
      snprintf(nm, sizeof nm, "bf16x3 [dbg %d]", dbg);
      run_chain<Mode<KIND_BF16, 3>>(p, nullptr, nullptr, 10, sms, h, o, dbg); pr(nm, h, o);
    }
  }
  printf(fails ? "RESULT: FAIL (%d)\n" : "RESULT: PASS\n", fails);
  return fails ? 1 : 0;
}
