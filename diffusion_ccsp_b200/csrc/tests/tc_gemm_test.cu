// tc_gemm_test.cu — standalone numerics + throughput check of the tcgen05 GEMM kernels (kernels_tc.cuh)
// against a double-precision CPU reference.  Built and run on the B200 box:
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tc_gemm_test tc_gemm_test.cu
//   timeout 120 ./tc_gemm_test [perf]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <random>
#include <vector>

#include "../kernels_tc.cuh"

namespace ccsp {
void set_error(const std::string &) {}
void count_launch() {}
}  // namespace ccsp

using namespace ccsp;
using namespace ccsp::tc;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(2);                                                                             \
    }                                                                                      \
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
  std::vector<float> S, tb;          // [rows][512], [groups][512]
  std::vector<float> Wd1, bd1, Wd2, bd2;   // [128][256], [128], [P][128], [P]
};

static Problem make_problem(int n_nodes, int m_tiles, int groups, unsigned seed) {
  Problem p;
  p.n_nodes = n_nodes; p.m_tiles = m_tiles; p.groups = groups; p.P = 4;
  std::mt19937 rng(seed);
  std::uniform_real_distribution<float> u(-1.f, 1.f);
  const int rows = m_tiles * 256;
  p.node_emb.resize((size_t)(n_nodes + 1) * 256);
  for (auto &v : p.node_emb) v = u(rng);
  for (int k = 0; k < 256; ++k) p.node_emb[(size_t)n_nodes * 256 + k] = 0.f;
  p.idx0.resize(rows); p.idx1.resize(rows);
  for (int r = 0; r < rows; ++r) { p.idx0[r] = rng() % (n_nodes + 1); p.idx1[r] = rng() % (n_nodes + 1); }
  p.tile_type.resize(m_tiles);
  for (int t = 0; t < m_tiles; ++t) p.tile_type[t] = t % groups;
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

template <class C>
static std::vector<uint8_t> pack_l1(const Problem &p) {
  const size_t per = (size_t)2 * (512 / C::KC) * C::B_STAGE;
  std::vector<uint8_t> blob(per * p.groups);
  for (int g = 0; g < p.groups; ++g) pack_b_blob<C>(&p.W[(size_t)g * 512 * 512], 512, 0, 512, 512, blob.data() + g * per);
  return blob;
}

struct Report { double max_err, max_ref, ms; };

template <class C>
static Report run_l1(const Problem &p, const std::vector<double> *ref, int iters, int num_sms, int dbg = 0) {
  const int rows = p.m_tiles * 256;
  float *d_emb = dev(p.node_emb), *d_S = dev(p.S), *d_tb = dev(p.tb), *d_H;
  int *d_i0 = dev(p.idx0), *d_i1 = dev(p.idx1), *d_tt = dev(p.tile_type);
  std::vector<uint8_t> blob = pack_l1<C>(p);
  uint8_t *d_blob = dev(blob);
  CK(cudaMalloc(&d_H, (size_t)rows * 512 * sizeof(float)));
  CK(cudaMemset(d_H, 0xFF, (size_t)rows * 512 * sizeof(float)));
  GemmArgs a;
  memset(&a, 0, sizeof(a));
  a.a_src[0] = d_emb; a.a_src[1] = d_emb; a.a_idx[0] = d_i0; a.a_idx[1] = d_i1; a.nseg = 2;
  a.b_blob = d_blob; a.tile_type = d_tt; a.num_m_tiles = p.m_tiles; a.n_tiles = 512 / C::NTILE;
  a.S = d_S; a.tb = d_tb; a.H = d_H; a.dbg = dbg;
  CK(launch_gemm_tc<C>(a, num_sms, 0));
  CK(cudaDeviceSynchronize());
  Report rep{0, 0, 0};
  if (ref) {
    std::vector<float> H((size_t)rows * 512);
    CK(cudaMemcpy(H.data(), d_H, H.size() * sizeof(float), cudaMemcpyDeviceToHost));
    int shown = 0;
    for (size_t i = 0; i < H.size(); ++i) {
      double e = fabs((double)H[i] - (*ref)[i]);
      if (!(e == e)) e = 1e30;
      if (e > rep.max_err) rep.max_err = e;
      if (fabs((*ref)[i]) > rep.max_ref) rep.max_ref = fabs((*ref)[i]);
      if (e > 0.05 && shown < 6) { printf("    mismatch row %zu col %zu: got %g want %g\n", i / 512, i % 512, H[i], (*ref)[i]); ++shown; }
    }
  }
  if (iters > 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) CK(launch_gemm_tc<C>(a, num_sms, 0));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    rep.ms = ms / iters;
  }
  cudaFree(d_emb); cudaFree(d_S); cudaFree(d_tb); cudaFree(d_H); cudaFree(d_i0); cudaFree(d_i1); cudaFree(d_tt); cudaFree(d_blob);
  return rep;
}

template <class C>
static Report run_dec(const Problem &p, const std::vector<float> &Hin, const std::vector<double> *ref, int iters, int num_sms) {
  const int rows = (int)(Hin.size() / 256);      // 2 * edge rows
  float *d_H = dev(Hin), *d_bd1 = dev(p.bd1), *d_w2 = dev(p.Wd2), *d_bd2 = dev(p.bd2), *d_o;
  std::vector<uint8_t> blob((size_t)(256 / C::KC) * C::B_STAGE);
  pack_b_blob<C>(p.Wd1.data(), 256, 0, 256, 128, blob.data());
  uint8_t *d_blob = dev(blob);
  CK(cudaMalloc(&d_o, (size_t)rows * p.P * sizeof(float)));
  CK(cudaMemset(d_o, 0xFF, (size_t)rows * p.P * sizeof(float)));
  GemmArgs a;
  memset(&a, 0, sizeof(a));
  a.a_src[0] = d_H; a.nseg = 1; a.b_blob = d_blob; a.num_m_tiles = rows / 256; a.n_tiles = 1;
  a.bd1 = d_bd1; a.Wd2 = d_w2; a.bd2 = d_bd2; a.P = p.P; a.o = d_o;
  CK(launch_gemm_tc<C>(a, num_sms, 0));
  CK(cudaDeviceSynchronize());
  Report rep{0, 0, 0};
  if (ref) {
    std::vector<float> o((size_t)rows * p.P);
    CK(cudaMemcpy(o.data(), d_o, o.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < o.size(); ++i) {
      double e = fabs((double)o[i] - (*ref)[i]);
      if (!(e == e)) e = 1e30;
      if (e > rep.max_err) rep.max_err = e;
      if (fabs((*ref)[i]) > rep.max_ref) rep.max_ref = fabs((*ref)[i]);
    }
  }
  if (iters > 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) CK(launch_gemm_tc<C>(a, num_sms, 0));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    rep.ms = ms / iters;
  }
  cudaFree(d_H); cudaFree(d_bd1); cudaFree(d_w2); cudaFree(d_bd2); cudaFree(d_o); cudaFree(d_blob);
  return rep;
}

int main(int argc, char **argv) {
  const bool perf = argc > 1 && !strcmp(argv[1], "perf");
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, smem/block optin %zu\n", prop.name, sms, prop.sharedMemPerBlockOptin);
  int fails = 0;
  {
    // ---- numerics: 11 tiles of 256 rows, 3 weight groups ---------------------------------------------
    Problem p = make_problem(1000, 11, 3, 1);
    const int rows = p.m_tiles * 256;
    std::vector<double> ref((size_t)rows * 512);
    std::vector<float> Href((size_t)rows * 512);
    for (int r = 0; r < rows; ++r) {
      const float *a0 = &p.node_emb[(size_t)p.idx0[r] * 256], *a1 = &p.node_emb[(size_t)p.idx1[r] * 256];
      const int g = p.tile_type[r / 256];
      for (int n = 0; n < 512; ++n) {
        const float *w = &p.W[((size_t)g * 512 + n) * 512];
        double acc = 0;
        for (int k = 0; k < 256; ++k) acc += (double)a0[k] * w[k] + (double)a1[k] * w[256 + k];
        double v = silu_d(acc + p.S[(size_t)r * 512 + n] + p.tb[g * 512 + n]);
        ref[(size_t)r * 512 + n] = v;
        Href[(size_t)r * 512 + n] = (float)v;
      }
    }
    std::vector<double> oref((size_t)rows * 2 * p.P);
    for (int q = 0; q < rows * 2; ++q) {
      double d[128];
      for (int j = 0; j < 128; ++j) {
        double acc = 0;
        for (int k = 0; k < 256; ++k) acc += (double)Href[(size_t)q * 256 + k] * p.Wd1[j * 256 + k];
        d[j] = silu_d(acc + p.bd1[j]);
      }
      for (int pp = 0; pp < p.P; ++pp) {
        double acc = 0;
        for (int j = 0; j < 128; ++j) acc += d[j] * p.Wd2[pp * 128 + j];
        oref[(size_t)q * p.P + pp] = acc + p.bd2[pp];
      }
    }
    auto chk = [&](const char *name, Report r, double tol) {
      bool ok = r.max_err <= tol * (r.max_ref > 1 ? r.max_ref : 1);
      printf("%-22s max_err %.3e (max|ref| %.3f) tol %.1e  %s\n", name, r.max_err, r.max_ref, tol, ok ? "PASS" : "FAIL");
      if (!ok) ++fails;
    };
    chk("l1  tf32x3", run_l1<Cfg<KIND_TF32, 3, 256, EPI_TC_L1>>(p, &ref, 0, sms), 3e-6);
    chk("l1  bf16x3", run_l1<Cfg<KIND_BF16, 3, 256, EPI_TC_L1>>(p, &ref, 0, sms), 1e-4);
    chk("l1  tf32", run_l1<Cfg<KIND_TF32, 1, 256, EPI_TC_L1>>(p, &ref, 0, sms), 5e-3);
    chk("l1  bf16", run_l1<Cfg<KIND_BF16, 1, 256, EPI_TC_L1>>(p, &ref, 0, sms), 4e-2);
    chk("dec tf32x3", run_dec<Cfg<KIND_TF32, 3, 128, EPI_TC_DEC>>(p, Href, &oref, 0, sms), 3e-6);
    chk("dec bf16x3", run_dec<Cfg<KIND_BF16, 3, 128, EPI_TC_DEC>>(p, Href, &oref, 0, sms), 1e-4);
    chk("dec tf32", run_dec<Cfg<KIND_TF32, 1, 128, EPI_TC_DEC>>(p, Href, &oref, 0, sms), 5e-3);
    chk("dec bf16", run_dec<Cfg<KIND_BF16, 1, 128, EPI_TC_DEC>>(p, Href, &oref, 0, sms), 4e-2);
  }
  if (perf && fails == 0) {
    // ---- throughput at the config-2 size: 317 edge tiles (81 152 rows), 13 weight groups ---------------
    Problem p = make_problem(9216, 317, 13, 2);
    const double fl1 = 2.0 * 317 * 256 * 512 * 512, fdec = 2.0 * 317 * 512 * 128 * 256;
    std::vector<float> Hin((size_t)317 * 256 * 512);
    for (size_t i = 0; i < Hin.size(); ++i) Hin[i] = (float)((i * 2654435761u) % 1000) / 1000.f - 0.5f;
    auto pr = [&](const char *name, Report r, double fl) { printf("%-34s %.3f ms  %.1f TFLOP/s (algorithmic)\n", name, r.ms, fl / r.ms / 1e9); };
    pr("l1  tf32x3", run_l1<Cfg<KIND_TF32, 3, 256, EPI_TC_L1>>(p, nullptr, 20, sms), fl1);
    pr("l1  bf16x3", run_l1<Cfg<KIND_BF16, 3, 256, EPI_TC_L1>>(p, nullptr, 20, sms), fl1);
    pr("l1  tf32", run_l1<Cfg<KIND_TF32, 1, 256, EPI_TC_L1>>(p, nullptr, 20, sms), fl1);
    pr("l1  bf16", run_l1<Cfg<KIND_BF16, 1, 256, EPI_TC_L1>>(p, nullptr, 20, sms), fl1);
    const char *abl[] = {"none", "noA", "noB", "noA,noB", "noEpiIO", "noA,noEpiIO", "noB,noEpiIO", "MMA+sync only",
                         "noMMA", "noMMA,noA", "noMMA,noB", "noMMA,noA,noB", "noMMA,noEpiIO", "", "", "sync skeleton"};
    for (int dbg : {1, 2, 3, 4, 7, 8, 9, 10, 12, 15}) {
      char nm[64];
      snprintf(nm, sizeof nm, "l1 tf32x3 [%s]", abl[dbg]);
      pr(nm, run_l1<Cfg<KIND_TF32, 3, 256, EPI_TC_L1>>(p, nullptr, 10, sms, dbg), fl1);
    }
    for (int dbg : {1, 2, 4, 7, 8}) {
      char nm[64];
      snprintf(nm, sizeof nm, "l1 bf16x3 [%s]", abl[dbg]);
      pr(nm, run_l1<Cfg<KIND_BF16, 3, 256, EPI_TC_L1>>(p, nullptr, 10, sms, dbg), fl1);
    }
    pr("dec tf32x3", run_dec<Cfg<KIND_TF32, 3, 128, EPI_TC_DEC>>(p, Hin, nullptr, 20, sms), fdec);
    pr("dec bf16x3", run_dec<Cfg<KIND_BF16, 3, 128, EPI_TC_DEC>>(p, Hin, nullptr, 20, sms), fdec);
    pr("dec tf32", run_dec<Cfg<KIND_TF32, 1, 128, EPI_TC_DEC>>(p, Hin, nullptr, 20, sms), fdec);
    pr("dec bf16", run_dec<Cfg<KIND_BF16, 1, 128, EPI_TC_DEC>>(p, Hin, nullptr, 20, sms), fdec);
  }
  printf(fails ? "RESULT: FAIL (%d)\n" : "RESULT: PASS\n", fails);
  return fails ? 1 : 0;
}
