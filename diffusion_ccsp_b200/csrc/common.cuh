// common.cuh — shared device helpers (activations, Philox, error plumbing) for libccsp_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#define CCSP_H 256            // hidden_dim
#define CCSP_H2 512           // 2 * hidden_dim (first-layer output width, one half per edge endpoint)
#define CCSP_HH 128           // hidden_dim / 2
#define CCSP_TILE_M 128       // edge rows per tensor-core tile
#define CCSP_CLUSTER 2        // thread-block-cluster size of the first-layer kernel (weight-stage multicast)
#define CCSP_PAD_M (CCSP_TILE_M * CCSP_CLUSTER)   // every constraint type is padded to whole cluster groups of tiles
#define CCSP_MAXP 8
#define CCSP_MAX_CHAINS 4    // independent scene groups of a plan that the pipelined sampling path interleaves

namespace ccsp {

// ---- error plumbing (host) ---------------------------------------------------------------------
void set_error(const std::string &msg);
void count_launch();

#define CCSP_CUDA_TRY(expr)                                                                          \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      ::ccsp::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                        ":" + std::to_string(__LINE__) + ")");                                      \
      return CCSP_ERR_CUDA;                                                                          \
    }                                                                                                \
  } while (0)

#define CCSP_LAUNCH_CHECK()                                                                          \
  do {                                                                                               \
    ::ccsp::count_launch();                                                                          \
    cudaError_t _e = cudaGetLastError();                                                             \
    if (_e != cudaSuccess) {                                                                         \
      ::ccsp::set_error(std::string("kernel launch failed: ") + cudaGetErrorString(_e) + " (" +      \
                        __FILE__ + ":" + std::to_string(__LINE__) + ")");                           \
      return CCSP_ERR_CUDA;                                                                          \
    }                                                                                                \
  } while (0)

// ---- blocked layout of the per-edge static term S [Epad, 512] ------------------------------------
// 32-row x 32-column blocks; inside a block: 8 pieces (4 floats each) x 32 rows x 16 B, so that a warp
// whose lanes own 32 consecutive rows reads/writes 512 contiguous bytes per instruction (the access
// pattern of a TMEM-lane-per-thread epilogue).  Returns the float offset of element (row, col).
__host__ __device__ __forceinline__ size_t blk_off(size_t row, int col) {
  return ((((row >> 5) * 16 + (size_t)(col >> 5)) * 8 + (size_t)((col & 31) >> 2)) * 32 + (row & 31)) * 4 + (col & 3);
}

// ---- activations (torch.nn.SiLU / torch.nn.Mish semantics, accurate libm versions) -------------
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ float mish_f(float x) {
  float sp = x > 20.0f ? x : log1pf(expf(x));   // torch softplus threshold = 20
  return x * tanhf(sp);
}

// ---- Philox4x32-10 (Salmon et al. 2011) + Box-Muller -------------------------------------------
struct Philox4 { uint32_t x, y, z, w; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  // rolled on purpose: the callers are latency-bound kernels that run this once per launch from a cold instruction
  // cache, where code size is time
#pragma unroll 1
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

// two standard normals from two 32-bit words
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float &n0, float &n1) {
  float u1 = ((float)a + 0.5f) * 2.3283064365386963e-10f;   // (0,1)
  float u2 = ((float)b + 0.5f) * 2.3283064365386963e-10f;
  float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

// P (<= 8) standard normals for (seed, draw, global node); identical for any sharding of the nodes
__device__ __forceinline__ void philox_normals(uint64_t seed, uint32_t draw, uint64_t node, int P, float *z) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    if (q * 4 < P) {
      Philox4 r = philox4x32_10((uint32_t)node, (uint32_t)(node >> 32), draw, (uint32_t)q,
                                (uint32_t)seed, (uint32_t)(seed >> 32));
      box_muller(r.x, r.y, z[q * 4 + 0], z[q * 4 + 1]);
      box_muller(r.z, r.w, z[q * 4 + 2], z[q * 4 + 3]);
    }
  }
}

}  // namespace ccsp
