// kernels_train.cuh — FP32 kernels of the training step: loss of GaussianDiffusion.p_losses
// (networks/ddpm.py:353-389) through ConstraintDiffuser.forward (networks/denoise_fn.py:453-537) and the gradient with
// respect to every parameter (the reference gets it from autograd: loss.backward(), ddpm.py:136-142, 533-534).
//
// Same restructuring as the sampling path (edges sorted by type so every type is one gathered GEMM, destination-CSR for a
// deterministic scatter, one time embedding per batch because the reference draws ONE t per batch, ddpm.py:388), but
// nothing is hoisted: geometry / grasp / time encoders are trainable, so every layer is evaluated and differentiated.
// Everything is true FP32 (FMA on CUDA cores, fixed summation order => bit-reproducible steps); the training step is not the
// headline path and at the reference's batch size (128 scenes ~ 10 k edges, ~40 GFLOP per step) a tiled SIMT GEMM keeps the
// whole step at a few milliseconds.  All heavy contractions go through ONE generic tiled kernel (k_sgemm) parameterised by
// accessor structs (gathered rows, per-tile weight group, split-K partials); the rest are small reductions.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "common.cuh"

namespace ccsp {
namespace train {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;
constexpr int TILE_ROWS = 64;              // every constraint type is padded to whole 64-row tiles
constexpr int MAX_TYPES = 16;
constexpr int MAX_SEG = 5;                 // [ (grasp_i) | geom_i | geom_j | pose_i | pose_j ]   (denoise_fn.py:346-354)

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float dsilu_f(float z) { const float s = sigmoid_f(z); return s * (1.0f + z * (1.0f - s)); }
__device__ __forceinline__ float dmish_f(float x) {
  const float sp = x > 20.0f ? x : log1pf(expf(x));
  const float th = tanhf(sp);
  const float dsp = x > 20.0f ? 1.0f : sigmoid_f(x);       // d softplus / dx (threshold 20 like torch)
  return th + x * (1.0f - th * th) * dsp;
}

// -------------------------------------------------------------------------------------------------------------
// generic tiled GEMM:  C(z; m, n) = sum_{k in [k0,k1)} A(z; m, k) * B(z; k, n)
//   P::shape(z, M, N, k0, k1);  P::a(z, m, k);  P::b(z, m0, k, n);  P::c(z, m, n, acc)
//   P::A_K_CONTIG / P::B_N_CONTIG pick the thread->element mapping of the tile loads so that global reads coalesce
// -------------------------------------------------------------------------------------------------------------
template <class P>
__global__ void __launch_bounds__(NT) k_sgemm(const P p) {
  const int z = blockIdx.z;
  int M, N, k0, k1;
  p.shape(z, M, N, k0, k1);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (m0 >= M || n0 >= N) return;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  for (int kk = k0; kk < k1; kk += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m, k;
      if (P::A_K_CONTIG) { k = tid & 15; m = (tid >> 4) + 16 * i; } else { m = tid & 63; k = (tid >> 6) + 4 * i; }
      As[k][m] = (m0 + m < M && kk + k < k1) ? p.a(z, m0 + m, kk + k) : 0.f;
      int n, kb;
      if (P::B_N_CONTIG) { n = tid & 63; kb = (tid >> 6) + 4 * i; } else { kb = tid & 15; n = (tid >> 4) + 16 * i; }
      Bs[kb][n] = (n0 + n < N && kk + kb < k1) ? p.b(z, m0, kk + kb, n0 + n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) p.c(z, m, n, acc[i][j]);
    }
}

// gathered first-layer input In(e, k): 256-wide segments, each a row of a node table picked through an index array
struct SegSrc {
  const float *base[MAX_SEG];
  const int *idx[MAX_SEG];
  __device__ __forceinline__ float at(int e, int k) const {
    const int s = k >> 8;
    return base[s][(size_t)idx[s][e] * CCSP_H + (k & 255)];
  }
};
struct PtrTable { const float *p[MAX_TYPES]; };
struct MutPtrTable { float *p[MAX_TYPES]; };

// ---- the GEMM instances ----------------------------------------------------------------------------------------
// (1) plain  Y[m, n] = act(sum_k X[m, k] W[n, k] + bias[n]),  W in nn.Linear layout [N, K]; stores pre-activation and activation
struct LinearFwd {
  static constexpr bool A_K_CONTIG = true, B_N_CONTIG = false;
  const float *X, *W, *bias;
  float *Z, *Y;
  int M_, N_, K_;
  __device__ void shape(int, int &M, int &N, int &k0, int &k1) const { M = M_; N = N_; k0 = 0; k1 = K_; }
  __device__ float a(int, int m, int k) const { return X[(size_t)m * K_ + k]; }
  __device__ float b(int, int, int k, int n) const { return W[(size_t)n * K_ + k]; }
  __device__ void c(int, int m, int n, float acc) const {
    const float z = acc + bias[n];
    Z[(size_t)m * N_ + n] = z;
    Y[(size_t)m * N_ + n] = silu_f(z);
  }
};

// (2) first layer:  Z[e, o] = sum_k In(e, k) W_c[o, k] + bias_c[o],  c = type of the row tile; bias_c holds b_c + W_c[:, time] temb
struct EdgeL1Fwd {
  static constexpr bool A_K_CONTIG = true, B_N_CONTIG = false;
  SegSrc in;
  PtrTable W;
  const int *tile_type;
  const float *bias;       // [C, 512]
  float *Z, *H;
  int Epad, Kseg, Kin;
  __device__ void shape(int, int &M, int &N, int &k0, int &k1) const { M = Epad; N = CCSP_H2; k0 = 0; k1 = Kseg; }
  __device__ float a(int, int m, int k) const { return in.at(m, k); }
  __device__ float b(int, int m0, int k, int n) const { return W.p[tile_type[m0 / TILE_ROWS]][(size_t)n * Kin + k]; }
  __device__ void c(int, int m, int n, float acc) const {
    const float z = acc + bias[tile_type[m / TILE_ROWS] * CCSP_H2 + n];
    Z[(size_t)m * CCSP_H2 + n] = z;
    H[(size_t)m * CCSP_H2 + n] = silu_f(z);
  }
};

// (3) dX[m, n] = (sum_k dY[m, k] W[k, n]) * silu'(Zx[m, n])      (back through a Linear into the previous activation)
struct LinearBwdInput {
  static constexpr bool A_K_CONTIG = true, B_N_CONTIG = true;
  const float *dY, *W, *Zx;
  float *dX;
  int M_, N_, K_;
  __device__ void shape(int, int &M, int &N, int &k0, int &k1) const { M = M_; N = N_; k0 = 0; k1 = K_; }
  __device__ float a(int, int m, int k) const { return dY[(size_t)m * K_ + k]; }
  __device__ float b(int, int, int k, int n) const { return W[(size_t)k * N_ + n]; }
  __device__ void c(int, int m, int n, float acc) const { dX[(size_t)m * N_ + n] = acc * dsilu_f(Zx[(size_t)m * N_ + n]); }
};

// (4) dW[m, n] = sum_r dY[r, m] X[r, n]  over rows r, split-K: slice z handles rows [z*chunk, (z+1)*chunk); partials [z][M][N]
struct LinearBwdWeight {
  static constexpr bool A_K_CONTIG = false, B_N_CONTIG = true;
  const float *dY, *X;
  float *part;
  int M_, N_, R_, chunk;
  __device__ void shape(int z, int &M, int &N, int &k0, int &k1) const {
    M = M_; N = N_; k0 = z * chunk; k1 = min(R_, k0 + chunk);
  }
  __device__ float a(int, int m, int r) const { return dY[(size_t)r * M_ + m]; }
  __device__ float b(int, int, int r, int n) const { return X[(size_t)r * N_ + n]; }
  __device__ void c(int z, int m, int n, float acc) const { part[((size_t)z * M_ + m) * N_ + n] = acc; }
};

// (5) dW_c[o, k] = sum_{e of type c} dZ[e, o] In(e, k)     grid.z = c; rows of type c = [start[c], start[c+1])
struct EdgeL1BwdWeight {
  static constexpr bool A_K_CONTIG = false, B_N_CONTIG = true;
  SegSrc in;
  const float *dZ;
  MutPtrTable dW;
  int start[MAX_TYPES + 1];
  int Kseg, Kin;
  __device__ void shape(int z, int &M, int &N, int &k0, int &k1) const { M = CCSP_H2; N = Kseg; k0 = start[z]; k1 = start[z + 1]; }
  __device__ float a(int, int m, int e) const { return dZ[(size_t)e * CCSP_H2 + m]; }
  __device__ float b(int, int, int e, int n) const { return in.at(e, n); }
  __device__ void c(int z, int m, int n, float acc) const { dW.p[z][(size_t)m * Kin + n] = acc; }
};

// (6) dIn[e, k] = sum_o dZ[e, o] W_c[o, k]
struct EdgeL1BwdInput {
  static constexpr bool A_K_CONTIG = true, B_N_CONTIG = true;
  const float *dZ;
  PtrTable W;
  const int *tile_type;
  float *dIn;
  int Epad, Kseg, Kin;     // Kseg = width of the column window written to dIn (row stride Kseg)
  int col0;                // first input column of the window (0: all gathered segments; energy form: the pose segments only)
  __device__ void shape(int, int &M, int &N, int &k0, int &k1) const { M = Epad; N = Kseg; k0 = 0; k1 = CCSP_H2; }
  __device__ float a(int, int m, int k) const { return dZ[(size_t)m * CCSP_H2 + k]; }
  __device__ float b(int, int m0, int k, int n) const { return W.p[tile_type[m0 / TILE_ROWS]][(size_t)k * Kin + col0 + n]; }
  __device__ void c(int, int m, int n, float acc) const { dIn[(size_t)m * Kseg + n] = acc; }
};

// sum split-K partials in slice order (deterministic):  out[i] = sum_z part[z][i]
__global__ void k_reduce_parts(const float *part, int slices, size_t count, float *out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float s = 0.f;
  for (int z = 0; z < slices; ++z) s += part[(size_t)z * count + i];
  out[i] = s;
}

// column sums over a row range, fixed order:  out[g][j] = sum_{r in [start[g], start[g+1])} M[r * ld + j]      (bias gradients)
// two stages so that long ranges are not one block's serial loop: grid (ceil(cols / 32), groups, slices) writes
// part[slice][g][j] (each block: 32 columns x 8 row lanes over its slice of the range), k_colsum_finish adds the slices in order
constexpr int COLSUM_SLICES = 32;
struct ColSumArgs {
  const float *M;
  int ld, cols, groups;
  int start[MAX_TYPES + 1];      // row range per group (grid.y)
  float *part;                   // [COLSUM_SLICES, groups, cols]
  float *out;                    // [groups, cols]
};
__global__ void __launch_bounds__(256) k_colsum(const ColSumArgs A) {
  __shared__ float red[8][33];
  const int g = blockIdx.y, z = blockIdx.z, col = blockIdx.x * 32 + (threadIdx.x & 31), lane_r = threadIdx.x >> 5;
  const int r0 = A.start[g], r1 = A.start[g + 1];
  const int chunk = (r1 - r0 + COLSUM_SLICES - 1) / COLSUM_SLICES;
  const int b = r0 + z * chunk, e = min(r1, b + chunk);
  float s = 0.f;
  if (col < A.cols)
    for (int r = b + lane_r; r < e; r += 8) s += A.M[(size_t)r * A.ld + col];
  red[lane_r][threadIdx.x & 31] = s;
  __syncthreads();
  if (lane_r == 0 && col < A.cols) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
    A.part[((size_t)z * A.groups + g) * A.cols + col] = t;
  }
}
__global__ void k_colsum_finish(const ColSumArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;       // (g, col)
  if (i >= A.groups * A.cols) return;
  float t = 0.f;
  for (int z = 0; z < COLSUM_SLICES; ++z) t += A.part[(size_t)z * A.groups * A.cols + i];
  A.out[i] = t;
}

// ---- encoders, first layer (K = G <= 8): z1 = X[:, off:off+G] W0^T + b0, a1 = silu(z1) ---------------------------
__global__ void k_enc1_fwd(const float *X, int ldx, int off, int G, int n, const float *W0, const float *b0, float *z1, float *a1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * CCSP_HH) return;
  const int r = i / CCSP_HH, j = i % CCSP_HH;
  float s = b0[j];
  for (int k = 0; k < G; ++k) s = fmaf(X[(size_t)r * ldx + off + k], W0[j * G + k], s);
  z1[i] = s;
  a1[i] = silu_f(s);
}
// dW0[j, k] = sum_r dz1[r, j] X[r, off + k];  db0[j] = sum_r dz1[r, j]        one block per j, fixed-order tree
__global__ void __launch_bounds__(256) k_enc1_bwd(const float *X, int ldx, int off, int G, int n, const float *dz1, float *dW0, float *db0) {
  __shared__ float red[256];
  const int j = blockIdx.x;
  float acc[CCSP_MAXP + 1];
  for (int k = 0; k <= G; ++k) acc[k] = 0.f;
  for (int r = threadIdx.x; r < n; r += 256) {
    const float d = dz1[(size_t)r * CCSP_HH + j];
    for (int k = 0; k < G; ++k) acc[k] = fmaf(d, X[(size_t)r * ldx + off + k], acc[k]);
    acc[G] += d;
  }
  for (int k = 0; k <= G; ++k) {
    red[threadIdx.x] = acc[k];
    __syncthreads();
    for (int s = 128; s; s >>= 1) {
      if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) { if (k < G) dW0[j * G + k] = red[0]; else db0[j] = red[0]; }
    __syncthreads();
  }
}

// ---- time MLP (denoise_fn.py:43-50, 259-264), one t per batch -------------------------------------------------------
// y[j] = act(b[j] + sum_k W[j, k] x[k]): one warp per output row (coalesced reads of the row, fixed-order tree), 8 rows per block.
// MODE 0: x = sinusoidal embedding of t (computed here, also stored to `emb`), act = mish, stores z and mish(z)
// MODE 1: x given, no activation
template <int MODE>
__global__ void __launch_bounds__(256) k_time_gemv(int t, const float *freqs, const float *xin, const float *W, const float *b, int N, int K,
                                                   float *emb, float *z, float *y) {
  extern __shared__ float sx[];
  if (MODE == 0) {
    for (int i = threadIdx.x; i < CCSP_H; i += 256) {
      const float arg = (float)t * freqs[i & (CCSP_HH - 1)];
      const float v = i < CCSP_HH ? sinf(arg) : cosf(arg);
      sx[i] = v;
      if (blockIdx.x == 0) emb[i] = v;
    }
  } else {
    for (int i = threadIdx.x; i < K; i += 256) sx[i] = xin[i];
  }
  __syncthreads();
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= N) return;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s = fmaf(W[(size_t)j * K + k], sx[k], s);
  for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) {
    s += b[j];
    if (MODE == 0) { z[j] = s; y[j] = mish_f(s); } else { y[j] = s; }
  }
}
// bias_c[o] = b_c[o] + sum_k W_c[o, tcol + k] temb[k]            grid (C), block 512
__global__ void __launch_bounds__(512) k_time_bias_fwd(PtrTable W, PtrTable b, int Kin, int tcol, const float *temb, float *bias) {
  __shared__ float st[CCSP_H];
  if (threadIdx.x < CCSP_H) st[threadIdx.x] = temb[threadIdx.x];
  __syncthreads();
  const int c = blockIdx.x, o = threadIdx.x;
  const float *w = W.p[c] + (size_t)o * Kin + tcol;
  float s = b.p[c][o];
  for (int k = 0; k < CCSP_H; ++k) s = fmaf(w[k], st[k], s);
  bias[c * CCSP_H2 + o] = s;
}
// time columns of dW_c and the gradient of temb:  dW_c[o, tcol + k] = db_c[o] temb[k];  dtemb[k] = sum_c sum_o W_c[o, tcol + k] db_c[o]
// grid (C, 8): block (c, s) handles output rows o in [64 s, 64 s + 64): writes its outer-product rows and the partial
// dtemb_part[c][s][k]; k_time_bwd adds the partials in (c, s) order => deterministic
__global__ void __launch_bounds__(256) k_time_cols_bwd(PtrTable W, MutPtrTable dW, int Kin, int tcol, const float *temb,
                                                        const float *dbias /*[C,512]*/, const int *type_rows, float *dtemb_part) {
  const int c = blockIdx.x, sl = blockIdx.y, k = threadIdx.x;
  const float tk = temb[k];
  float s = 0.f;
  const bool live = type_rows[c] != 0;                     // types without edges are not part of the graph (denoise_fn.py:514-515)
  for (int o = sl * 64; o < sl * 64 + 64; ++o) {
    const float d = dbias[c * CCSP_H2 + o];
    dW.p[c][(size_t)o * Kin + tcol + k] = d * tk;
    if (live) s = fmaf(W.p[c][(size_t)o * Kin + tcol + k], d, s);
  }
  dtemb_part[((size_t)c * 8 + sl) * CCSP_H + k] = s;
}
// back through the time MLP.  grid (64), block 256: every block recomputes dtemb (C*8 partials) and dz1 = (W3^T dtemb) * mish'(z1)
// for its 16 hidden units, then writes its slices of the two outer products
__global__ void __launch_bounds__(256) k_time_bwd(const float *dtemb_part, int nparts, const float *emb, const float *z1, const float *a1,
                                                  const float *W3, float *dW1, float *db1, float *dW3, float *db3) {
  __shared__ float sd[CCSP_H], se[CCSP_H], sdz[16], sa[16];
  const int tid = threadIdx.x, j0 = blockIdx.x * 16;
  {
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += dtemb_part[(size_t)p * CCSP_H + tid];
    sd[tid] = s; se[tid] = emb[tid];
    if (blockIdx.x == 0) db3[tid] = s;
  }
  __syncthreads();
  {  // 16 hidden units of this block: 16 lanes per unit over k, fixed-order tree
    const int u = tid >> 4, l = tid & 15, j = j0 + u;
    float s = 0.f;
    for (int k = l; k < CCSP_H; k += 16) s = fmaf(W3[(size_t)k * 4 * CCSP_H + j], sd[k], s);
    for (int off = 8; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off, 16);
    if (l == 0) {
      const float dz = s * dmish_f(z1[j]);
      sdz[u] = dz; db1[j] = dz; sa[u] = a1[j];
    }
  }
  __syncthreads();
  for (int i = tid; i < CCSP_H * 16; i += 256) {            // dW3[k, j0 + u] = dtemb[k] a1[j]
    const int k = i >> 4, u = i & 15;
    dW3[(size_t)k * 4 * CCSP_H + j0 + u] = sd[k] * sa[u];
  }
  for (int i = tid; i < 16 * CCSP_H; i += 256) {            // dW1[j0 + u, k] = dz1[j] emb[k]
    const int u = i >> 8, k = i & 255;
    dW1[(size_t)(j0 + u) * CCSP_H + k] = sdz[u] * se[k];
  }
}

// ---- q_sample (ddpm.py:353-361): x_t = a x_0 + b noise, pinned rows keep x_0; the noise is used as given (the caller zeroes the
// pinned rows when it draws it itself, conditional_noise ddpm.py:114-117) ---------------------------------------------------
__global__ void k_q_sample(const float *x0, const float *noise_in, const signed char *mask, int n, int P, float sa, float sb,
                           float *noise, float *xt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * P) return;
  const bool m = mask[i / P] != 0;
  const float z = noise_in[i];
  noise[i] = z;
  xt[i] = m ? x0[i] : __fadd_rn(__fmul_rn(sa, x0[i]), __fmul_rn(sb, z));
}

// ---- decoder second layer (N = P <= 8):  O[r, p] = bd2[p] + sum_k A1[r, k] Wd2[p, k] ----------------------------------------
__global__ void __launch_bounds__(256) k_dec2_fwd(const float *A1, const float *Wd2, const float *bd2, int rows, int P, float *O) {
  __shared__ float sw[CCSP_MAXP * CCSP_HH];
  for (int i = threadIdx.x; i < P * CCSP_HH; i += 256) sw[i] = Wd2[i];
  __syncthreads();
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= rows) return;
  float acc[CCSP_MAXP];
  for (int p = 0; p < P; ++p) acc[p] = bd2[p];
  const float4 *a4 = reinterpret_cast<const float4 *>(A1 + (size_t)r * CCSP_HH);
  for (int k4 = 0; k4 < CCSP_HH / 4; ++k4) {
    const float4 a = a4[k4];
    for (int p = 0; p < P; ++p) {
      const float *w = sw + p * CCSP_HH + k4 * 4;
      acc[p] = fmaf(a.x, w[0], fmaf(a.y, w[1], fmaf(a.z, w[2], fmaf(a.w, w[3], acc[p]))));
    }
  }
  for (int p = 0; p < P; ++p) O[(size_t)r * P + p] = acc[p];
}

// ---- scatter-reduce, normalise, pin, loss terms, dOut   (denoise_fn.py:377-389, 523-533; ddpm.py:379-384) ------------------
// one thread per node; err[v] = sum_p loss term;  dOut[v, p] = scale * dloss/dout (0 for pinned rows)
__global__ void k_node_loss(const float *O, const int *node_ptr, const int *node_src, const signed char *mask, const float *xtail,
                            const float *noise, int n, int P, int normalize, int loss_l1, float inv_count, float scale,
                            float *out, float *err, float *dOut) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  float acc[CCSP_MAXP];
  for (int p = 0; p < P; ++p) acc[p] = 0.f;
  const int b = node_ptr[v], e = node_ptr[v + 1];
  for (int i = b; i < e; ++i) {
    const float *row = O + (size_t)node_src[i] * P;
    for (int p = 0; p < P; ++p) acc[p] += row[p];
  }
  const float nrm = normalize ? sqrtf((float)(e - b)) : 1.0f;
  const bool m = mask[v] != 0;
  float es = 0.f;
  for (int p = 0; p < P; ++p) {
    const float o = m ? xtail[(size_t)v * P + p] : acc[p] / nrm;
    out[(size_t)v * P + p] = o;
    const float d = o - noise[(size_t)v * P + p];
    es += loss_l1 ? fabsf(d) : d * d;
    const float g = loss_l1 ? (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) : 2.0f * d;
    dOut[(size_t)v * P + p] = m ? 0.f : scale * g * inv_count / nrm;       // already divided by sqrt(deg): gradient of the row sum
  }
  err[v] = es;
}
// loss = (sum_v err[v]) / (n P), fixed order; single block
__global__ void __launch_bounds__(256) k_loss_reduce(const float *err, int n, float inv_count, float *loss) {
  __shared__ float red[256];
  float s = 0.f;
  for (int v = threadIdx.x; v < n; v += 256) s += err[v];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = red[0] * inv_count;
}

// ---- dO rows (gather of dOut through the edge endpoints), then dD1 = (dO Wd2) * silu'(D1) ------------------------------------
__global__ void k_dO(const float *dOut, const int *src_i, const int *src_j, int n, int rows, int P, float *dO) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * P) return;
  const int r = i / P, p = i % P;
  const int node = (r & 1) ? src_j[r >> 1] : src_i[r >> 1];
  dO[i] = node < n ? dOut[(size_t)node * P + p] : 0.f;        // padded rows feed nothing
}
__global__ void __launch_bounds__(256) k_dD1(const float *dO, const float *Wd2, const float *D1, int rows, int P, float *dD1) {
  __shared__ float sw[CCSP_MAXP * CCSP_HH];
  for (int i = threadIdx.x; i < P * CCSP_HH; i += 256) sw[i] = Wd2[i];
  __syncthreads();
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= (size_t)rows * CCSP_HH) return;
  const int r = (int)(i / CCSP_HH), k = (int)(i % CCSP_HH);
  float s = 0.f;
  for (int p = 0; p < P; ++p) s = fmaf(dO[(size_t)r * P + p], sw[p * CCSP_HH + k], s);
  dD1[i] = s * dsilu_f(D1[i]);
}

// ---- node-side reduction of dIn: gradient w.r.t. the node embeddings, times silu'(z2) ---------------------------------------
// one block per node, one thread per embedding column; incident (edge, slot) rows in CSR order => deterministic
// seg_of[table][slot] = first-layer segment that holds this table's row for this endpoint slot (-1: none)
struct NodeBwdArgs {
  const float *dIn;
  const int *node_ptr, *node_src;
  int Kseg, ntab;
  int seg_of[3][2];
  const float *z2[3];          // pre-activations of the second encoder layer per table [n+1, 256]
  float *dz2[3];               // out: d loss / d z2
};
__global__ void __launch_bounds__(256) k_node_bwd(const NodeBwdArgs A) {
  const int v = blockIdx.x, col = threadIdx.x;
  const int b = A.node_ptr[v], e = A.node_ptr[v + 1];
  float acc[3] = {0.f, 0.f, 0.f};
  for (int i = b; i < e; ++i) {
    const int row = A.node_src[i] >> 1, slot = A.node_src[i] & 1;
    for (int t = 0; t < A.ntab; ++t) {
      const int seg = A.seg_of[t][slot];
      if (seg >= 0) acc[t] += A.dIn[(size_t)row * A.Kseg + seg * CCSP_H + col];
    }
  }
  for (int t = 0; t < A.ntab; ++t) {
    const size_t o = (size_t)v * CCSP_H + col;
    A.dz2[t][o] = acc[t] * dsilu_f(A.z2[t][o]);
  }
}

// ---- energy form (denoise_fn.py:373-375, 518-521, 539-548): E = sum over edges and both endpoints of |o - x[arg]|^2 ---------------
// dO[r, :] = dE/do = 2 (o - x[node(r)]);  err[r] = |o - x|^2;  padded rows contribute nothing
__global__ void k_energy_dO(const float *O, const float *x, const int *src_i, const int *src_j, int n, int rows, int P, float *dO,
                            float *err) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int node = (r & 1) ? src_j[r >> 1] : src_i[r >> 1];
  float e = 0.f;
  for (int p = 0; p < P; ++p) {
    const float d = node < n ? O[(size_t)r * P + p] - x[(size_t)node * P + p] : 0.f;
    dO[(size_t)r * P + p] = 2.0f * d;
    e += d * d;
  }
  err[r] = e;
}
// dE/dx[v, p] = sum_j dz1[v, j] W0[j, p]   (through the pose encoder)   -  sum_{rows at v} dO[row, p]   (the direct -x term)
__global__ void k_energy_grad(const float *dz1, const float *W0, const float *dO, const int *node_ptr, const int *node_src, int n, int P,
                              float *grad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * P) return;
  const int v = i / P, p = i % P;
  float s = 0.f;
  for (int j = 0; j < CCSP_HH; ++j) s = fmaf(dz1[(size_t)v * CCSP_HH + j], W0[j * P + p], s);
  float d = 0.f;
  for (int k = node_ptr[v]; k < node_ptr[v + 1]; ++k) d += dO[(size_t)node_src[k] * P + p];
  grad[i] = s - d;
}

// ---- Adam (torch.optim.Adam defaults: no weight decay, no amsgrad; ddpm.py:466) ------------------------------------------------
__global__ void k_adam(float *p, const float *g, float *m, float *v, size_t count, float lr, float b1, float b2, float eps,
                       float bc1, float bc2_sqrt) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float gi = g[i];
  const float mi = m[i] + (gi - m[i]) * (1.0f - b1);             // exp_avg.lerp_(grad, 1 - beta1)
  const float vi = v[i] * b2 + (1.0f - b2) * gi * gi;            // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  m[i] = mi; v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = p[i] - (lr / bc1) * (mi / denom);
}

}  // namespace train
}  // namespace ccsp
