// kernels_simt.cuh — FP32 CUDA-core kernels.
//
//  * run-constant precompute (plan/model build): encoders, time-embedding table, per-edge static term;
//  * the node kernel (deterministic scatter-reduce + DDPM/ULA update + noise + pin + pose encoder),
//    which is shared by every CcspMath mode;
//  * the FP32 validation path for the two dense per-edge layers (CCSP_MATH_FP32).
//
// Reference semantics cited per kernel (paths relative to the reference checkout).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace ccsp {

// =================================================================================================
// Two-layer encoder  Linear(Din,128) SiLU Linear(128,256) SiLU      (denoise_fn.py:227-250)
// Used at plan build for geom_encoder / grasp_encoder (run-constant inputs).
//   x: [n, ldx] rows, columns [col0, col0+Din);  out: [n_out, 256] with rows >= n zero-filled
//   W0: [128, Din] (reference layout), W2t: [128][256] (transposed), 16 rows per block.
// =================================================================================================
#define ENC_ROWS 16

__device__ __forceinline__ void encoder_rows_16(const float (*xin)[CCSP_MAXP], int Din,
                                                const float *__restrict__ W0, const float *__restrict__ b0,
                                                const float *__restrict__ W2t, const float *__restrict__ b2,
                                                float (*h)[CCSP_HH], float *acc_out /*[ENC_ROWS]*/) {
  const int tid = threadIdx.x;
  // layer 1: 16 x 128 outputs over 256 threads
  for (int idx = tid; idx < ENC_ROWS * CCSP_HH; idx += 256) {
    int r = idx / CCSP_HH, j = idx % CCSP_HH;
    float acc = 0.f;
    for (int d = 0; d < Din; ++d) acc = fmaf(xin[r][d], __ldg(&W0[j * Din + d]), acc);
    h[r][j] = silu_f(acc + __ldg(&b0[j]));
  }
  __syncthreads();
  // layer 2: thread j owns output column j for all 16 rows
  float acc[ENC_ROWS];
#pragma unroll
  for (int r = 0; r < ENC_ROWS; ++r) acc[r] = 0.f;
  for (int k = 0; k < CCSP_HH; ++k) {
    float w = __ldg(&W2t[k * CCSP_H + tid]);
#pragma unroll
    for (int r = 0; r < ENC_ROWS; ++r) acc[r] = fmaf(h[r][k], w, acc[r]);
  }
  float bj = __ldg(&b2[tid]);
#pragma unroll
  for (int r = 0; r < ENC_ROWS; ++r) acc_out[r] = silu_f(acc[r] + bj);
}

__global__ void __launch_bounds__(256)
k_encode_rows(const float *__restrict__ x, int ldx, int col0, int Din, int n, int n_out,
              const float *__restrict__ W0, const float *__restrict__ b0,
              const float *__restrict__ W2t, const float *__restrict__ b2, float *__restrict__ out) {
  __shared__ float xin[ENC_ROWS][CCSP_MAXP];
  __shared__ float h[ENC_ROWS][CCSP_HH];
  const int tid = threadIdx.x, row0 = blockIdx.x * ENC_ROWS;
  if (tid < ENC_ROWS * CCSP_MAXP) {
    int r = tid / CCSP_MAXP, d = tid % CCSP_MAXP, row = row0 + r;
    xin[r][d] = (row < n && d < Din) ? x[(size_t)row * ldx + col0 + d] : 0.f;
  }
  __syncthreads();
  float o[ENC_ROWS];
  encoder_rows_16(xin, Din, W0, b0, W2t, b2, h, o);
#pragma unroll
  for (int r = 0; r < ENC_ROWS; ++r) {
    int row = row0 + r;
    if (row < n_out) out[(size_t)row * CCSP_H + tid] = row < n ? o[r] : 0.f;
  }
}

// =================================================================================================
// Time embedding table  time_mlp(t) for t in [0,T)           (denoise_fn.py:43-50, 259-264)
// The reference evaluates this MLP on E_c identical rows per type per call; it depends on t only.
//   freqs: [128] host-computed exp(k * -ln(1e4)/127) in FP32; W1t [256][1024], W3t [1024][256]
// =================================================================================================
__global__ void __launch_bounds__(256)
k_time_embed(const float *__restrict__ freqs, const float *__restrict__ W1t, const float *__restrict__ b1,
             const float *__restrict__ W3t, const float *__restrict__ b3, float *__restrict__ temb /*[T,256]*/) {
  __shared__ float pos[CCSP_H];
  __shared__ float hid[4 * CCSP_H];
  const int tid = threadIdx.x, t = blockIdx.x;
  if (tid < CCSP_HH) {
    float a = (float)t * freqs[tid];
    pos[tid] = sinf(a);
    pos[tid + CCSP_HH] = cosf(a);
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    int col = q * 256 + tid;
    float acc = 0.f;
    for (int k = 0; k < CCSP_H; ++k) acc = fmaf(pos[k], __ldg(&W1t[k * 1024 + col]), acc);
    hid[col] = mish_f(acc + __ldg(&b1[col]));
  }
  __syncthreads();
  float acc = 0.f;
  for (int k = 0; k < 4 * CCSP_H; ++k) acc = fmaf(hid[k], __ldg(&W3t[k * CCSP_H + tid]), acc);
  temb[(size_t)t * CCSP_H + tid] = acc + __ldg(&b3[tid]);
}

// Per-(t, type) time term of the first layer: tb[t][c][:] = mlps[c].weight[:, time cols] @ temb[t]
//   Wtt: [C][256][512] (transposed time columns).  grid (T, C).
__global__ void __launch_bounds__(256)
k_time_bias(const float *__restrict__ temb, const float *__restrict__ Wtt, int C, float *__restrict__ tb) {
  __shared__ float te[CCSP_H];
  const int tid = threadIdx.x, t = blockIdx.x, c = blockIdx.y;
  te[tid] = temb[(size_t)t * CCSP_H + tid];
  __syncthreads();
  const float *W = Wtt + (size_t)c * CCSP_H * CCSP_H2;
  float a0 = 0.f, a1 = 0.f;
  for (int k = 0; k < CCSP_H; ++k) {
    a0 = fmaf(te[k], __ldg(&W[k * CCSP_H2 + tid]), a0);
    a1 = fmaf(te[k], __ldg(&W[k * CCSP_H2 + 256 + tid]), a1);
  }
  float *o = tb + ((size_t)t * C + c) * CCSP_H2;
  o[tid] = a0;
  o[256 + tid] = a1;
}

// =================================================================================================
// Node kernel: deterministic scatter-reduce + update + pose encoder.
//
//   eps[v] = ( sum over incident (edge, slot) in reference order of o[edge, slot, :] ) / sqrt(deg[v])
//            (denoise_fn.py:377-389, 523-524; order = type-major, edge order, arg1 before arg2)
//   masked rows: eps[v] = x[v, -P:]                              (denoise_fn.py:532-533)
//   DDPM  (ddpm.py:230-258): x0 = a x - b eps; mean = c1 x0 + c2 x; x' = mean + sigma z
//   ULA   (ddpm.py:279-283, 960-964): x' = x + (-eps c_t) ss + z sqrt(2 ss)
//   pin   (ddpm.py:274, 334): x'[mask] = gt[mask]   (only at the end of a timestep / at init)
//   then  pe[v] = pose_encoder(x'[v])                            (denoise_fn.py:483)
// =================================================================================================
enum NodeMode { NODE_INIT = 0, NODE_DDPM = 1, NODE_ULA = 2, NODE_EPS_OUT = 3, NODE_ENCODE = 4 };

struct NodeArgs {
  int mode;
  int n, P, normalize;
  float a, b, c1, c2, sigma;     // DDPM coefficients at t (sigma = [t != 0] * exp(0.5 * logvar_t))
  float gscale, ss, std;         // ULA: c_t, step size, sqrt(2 ss)
  int pin;                       // re-pin masked rows after the update
  int has_xinit;                 // NODE_INIT: take x as given instead of 0.5 * z
  const float *z;                // injected draw [n,P] or nullptr -> Philox
  unsigned long long seed, node_offset;
  unsigned int draw;
  float *x;                      // state [n,P] (in/out)
  const float *x_in;             // NODE_ENCODE: poses to encode; NODE_INIT+has_xinit: start state
  float *eps_out;                // NODE_EPS_OUT: [n,P]
  float *hist;                   // history slot [n,P] or nullptr
  const float *o;                // per-edge decoder outputs, [E', 2, P]
  const int *node_ptr;           // [n+1]
  const int *node_src;           // [2E] -> row index into o viewed as [2E', P]
  const signed char *mask;       // [n]
  const float *gt, *xtail;       // [n,P]
  const float *W0, *b0, *W2t, *b2;   // pose encoder
  int pe_fmt;                    // 0: FP32 [n+1,256]; 1: TF32 split, 2: BF16 split  ([n+1][256 hi | 256 lo], kernels_tc.cuh)
  void *pe;
  long long *trace;              // developer aid (CCSP_NODE_TRACE=1): clock64 at the phase boundaries of CTA 0, else nullptr
  // persistent mode (k_node_tc<.., PERSIST = true>): the kernel runs ALL num_iters node iterations of a sample() (the init
  // step + one update per denoiser evaluation) next to the persistent edge kernel; per-iteration arguments come from `sched`
  const struct NodeEval *sched;  // [num_iters]
  int num_iters;
  size_t nP;                     // n * P: stride of the injected draws (z + draw * nP) and of the history slots
  unsigned *node_done;           // += 1 per CTA and iteration (x and pe of the iteration are written)
  unsigned *edge_done;           // iteration i >= 1 may start once it reaches i * edge_ctas
  unsigned edge_ctas;
  // pipelined chains (persistent mode): chain c owns node rows chain_row0[c] .. chain_row0[c + 1] (whole scenes) and the flag
  // words node_done + 32 c / edge_done + 32 c; a CTA walks the 64-row blocks blockIdx.x, + gridDim.x, .. of every chain in
  // turn, so a few node CTAs serve all nodes while the edge kernel works on the other chain.  num_chains <= 1: all rows.
  int num_chains;
  int chain_row0[CCSP_MAX_CHAINS + 1];
  unsigned chain_units[CCSP_MAX_CHAINS];   // edge units of chain c: iteration i may start at edge_done[c] >= i * 2 * chain_units[c]
  int node_partition;            // 1: CTA b serves chain b % num_chains only (small shards: one block per CTA, the chains' node phases
                                 // do not queue behind each other); 0: every CTA walks all chains in turn
};

// per-iteration arguments of the persistent node kernel (what ccsp_sample passes per launch otherwise)
struct NodeEval {
  int mode, pin;
  float a, b, c1, c2, sigma, gscale, ss, std;
  unsigned int draw;
  int hist_slot;                 // history slot this iteration writes, or -1
};

// 64 nodes per block, 512 threads.  Stage 0: one (node, component) pair per thread; stage 1: layer 1 of the
// pose encoder (P -> 128); stage 2: layer 2 (128 -> 256) as a register-tiled FP32 GEMM (4 rows x 8 columns
// per thread, W2t streamed through shared memory in 16-row slabs).  The tensor-core modes use the hardware
// exp2/rcp SiLU (as the edge kernels do); the FP32 validation mode keeps the libm-accurate one.
#define NODE_ROWS 64
#define NODE_THREADS 512
#define NODE_LDH (CCSP_HH + 4)
constexpr int NODE_SMEM_BYTES = (NODE_ROWS * CCSP_MAXP + NODE_ROWS * NODE_LDH + 2 * 16 * CCSP_H + CCSP_HH * (CCSP_MAXP + 1)) * 4;

__device__ __forceinline__ float silu_sel(float x, bool fast) { return fast ? __fdividef(x, 1.0f + __expf(-x)) : silu_f(x); }

__global__ void __launch_bounds__(NODE_THREADS) k_node(NodeArgs A) {
  extern __shared__ __align__(16) float node_smem[];
  float (*xs)[CCSP_MAXP] = reinterpret_cast<float (*)[CCSP_MAXP]>(node_smem);
  float (*h)[NODE_LDH] = reinterpret_cast<float (*)[NODE_LDH]>(node_smem + NODE_ROWS * CCSP_MAXP);
  float (*Bs)[16][CCSP_H] = reinterpret_cast<float (*)[16][CCSP_H]>(node_smem + NODE_ROWS * CCSP_MAXP + NODE_ROWS * NODE_LDH);
  float *w0s = node_smem + NODE_ROWS * CCSP_MAXP + NODE_ROWS * NODE_LDH + 2 * 16 * CCSP_H;   // [128][P] then b0 [128]
  const int tid = threadIdx.x, row0 = blockIdx.x * NODE_ROWS;
  const int P = A.P;
  const bool fast = A.pe_fmt != 0;
  if (A.mode != NODE_EPS_OUT) {
    for (int i = tid; i < CCSP_HH * P; i += NODE_THREADS) w0s[i] = __ldg(&A.W0[i]);
    for (int i = tid; i < CCSP_HH; i += NODE_THREADS) w0s[CCSP_HH * CCSP_MAXP + i] = __ldg(&A.b0[i]);
  }
  for (int slot = tid; slot < NODE_ROWS * CCSP_MAXP; slot += NODE_THREADS) {
    const int r = slot / CCSP_MAXP, p = slot % CCSP_MAXP, v = row0 + r;
    float xn = 0.f;
    if (v < A.n && p < P) {
      const size_t ix = (size_t)v * P + p;
      const bool masked = A.mask[v] != 0;
      if (A.mode == NODE_ENCODE) {
        xn = A.x_in[ix];
      } else {
        float zv = 0.f;
        const bool need_z = (A.mode == NODE_DDPM || A.mode == NODE_ULA || (A.mode == NODE_INIT && !A.has_xinit));
        if (need_z) {
          if (A.z) {
            zv = A.z[ix];
          } else {
            float zz[CCSP_MAXP];
            philox_normals(A.seed, A.draw, A.node_offset + (unsigned long long)v, P, zz);
            zv = zz[p];
          }
        }
        float eps = 0.f;
        if (A.mode != NODE_INIT) {
          if (masked) {
            eps = A.xtail[ix];
          } else {
            // same accumulation order as the reference's scatter_add_ (sequential adds); loads batched by 4
            const int k0 = A.node_ptr[v], k1 = A.node_ptr[v + 1];
            float acc = 0.f;
            int k = k0;
            for (; k + 4 <= k1; k += 4) {
              const int s0 = A.node_src[k], s1 = A.node_src[k + 1], s2 = A.node_src[k + 2], s3 = A.node_src[k + 3];
              const float o0 = A.o[(size_t)s0 * P + p], o1 = A.o[(size_t)s1 * P + p], o2 = A.o[(size_t)s2 * P + p], o3 = A.o[(size_t)s3 * P + p];
              acc = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc, o0), o1), o2), o3);
            }
            for (; k < k1; ++k) acc = __fadd_rn(acc, A.o[(size_t)A.node_src[k] * P + p]);
            eps = A.normalize ? acc / sqrtf((float)(k1 - k0)) : acc;
          }
        }
        if (A.mode == NODE_EPS_OUT) {
          A.eps_out[ix] = eps;
          xn = 0.f;
        } else {
          const float xv = (A.mode == NODE_INIT) ? 0.f : A.x[ix];
          if (A.mode == NODE_INIT) {
            xn = A.has_xinit ? A.x_in[ix] : __fmul_rn(0.5f, zv);
          } else if (A.mode == NODE_DDPM) {
            float x0 = __fsub_rn(__fmul_rn(A.a, xv), __fmul_rn(A.b, eps));
            float mean = __fadd_rn(__fmul_rn(A.c1, x0), __fmul_rn(A.c2, xv));
            xn = __fadd_rn(mean, __fmul_rn(A.sigma, zv));
          } else {  // NODE_ULA
            float grad = __fmul_rn(-eps, A.gscale);
            xn = __fadd_rn(__fadd_rn(xv, __fmul_rn(grad, A.ss)), __fmul_rn(zv, A.std));
          }
          if (A.pin && masked) xn = A.gt[ix];
          A.x[ix] = xn;
          if (A.hist) A.hist[ix] = xn;
        }
      }
    }
    xs[r][p] = xn;
  }
  if (A.mode == NODE_EPS_OUT) return;
  __syncthreads();
  // ---- layer 1: 64 x 128 outputs
  for (int idx = tid; idx < NODE_ROWS * CCSP_HH; idx += NODE_THREADS) {
    const int r = idx / CCSP_HH, j = idx % CCSP_HH;
    float acc = 0.f;
    for (int d = 0; d < P; ++d) acc = fmaf(xs[r][d], w0s[j * P + d], acc);
    h[r][j] = silu_sel(acc + w0s[CCSP_HH * CCSP_MAXP + j], fast);
  }
  // ---- layer 2: out[64 x 256] = h[64 x 128] . W2t[128 x 256]; thread = rows 4 ty.., cols 4 tx.. and 128 + 4 tx..
  const int tx = tid & 31, ty = tid >> 5;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  auto load_slab = [&](int slab, int buf) {      // 16 x 256 floats = 1024 float4, 2 per thread
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int f4 = q * NODE_THREADS + tid, kk = f4 >> 6, c4 = f4 & 63;
      *reinterpret_cast<float4 *>(&Bs[buf][kk][c4 * 4]) = __ldg(reinterpret_cast<const float4 *>(A.W2t + (size_t)(slab * 16 + kk) * CCSP_H + c4 * 4));
    }
  };
  load_slab(0, 0);
  __syncthreads();
  for (int slab = 0; slab < CCSP_HH / 16; ++slab) {
    const int buf = slab & 1;
    if (slab + 1 < CCSP_HH / 16) load_slab(slab + 1, buf ^ 1);
#pragma unroll
    for (int k4 = 0; k4 < 16; k4 += 4) {
      float4 a4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a4[i] = *reinterpret_cast<const float4 *>(&h[ty * 4 + i][slab * 16 + k4]);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][k4 + kk][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][k4 + kk][128 + tx * 4]);
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = kk == 0 ? a4[i].x : kk == 1 ? a4[i].y : kk == 2 ? a4[i].z : a4[i].w;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a, b[j], acc[i][j]);
        }
      }
    }
    __syncthreads();
  }
  // ---- epilogue: SiLU, write the pose embedding in the format the edge kernels consume
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int v = row0 + ty * 4 + i;
    if (v > A.n) continue;
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
      const int col = hb * 128 + tx * 4;
      const float4 bb = __ldg(reinterpret_cast<const float4 *>(A.b2 + col));
      float e[4] = {silu_sel(acc[i][hb * 4 + 0] + bb.x, fast), silu_sel(acc[i][hb * 4 + 1] + bb.y, fast),
                    silu_sel(acc[i][hb * 4 + 2] + bb.z, fast), silu_sel(acc[i][hb * 4 + 3] + bb.w, fast)};
      if (v == A.n) e[0] = e[1] = e[2] = e[3] = 0.f;           // zero row read by padded edges
      if (A.pe_fmt == 0) {
        *reinterpret_cast<float4 *>(reinterpret_cast<float *>(A.pe) + (size_t)v * CCSP_H + col) = make_float4(e[0], e[1], e[2], e[3]);
      } else if (A.pe_fmt == 1) {        // hi = rna_tf32(e), lo = rna_tf32(e - hi)
        uint4 hi, lo;
        uint32_t *hp = &hi.x, *lp = &lo.x;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hp[c]) : "f"(e[c]));
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lp[c]) : "f"(e[c] - __uint_as_float(hp[c])));
        }
        uint32_t *row = reinterpret_cast<uint32_t *>(A.pe) + (size_t)v * (2 * CCSP_H);
        *reinterpret_cast<uint4 *>(row + col) = hi;
        *reinterpret_cast<uint4 *>(row + CCSP_H + col) = lo;
      } else {                           // hi = rn_bf16(e), lo = rn_bf16(e - hi)
        float g[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) g[c] = __bfloat162float(__float2bfloat16_rn(e[c]));
        __nv_bfloat162 h01 = __floats2bfloat162_rn(g[0], g[1]), h23 = __floats2bfloat162_rn(g[2], g[3]);
        __nv_bfloat162 l01 = __floats2bfloat162_rn(e[0] - g[0], e[1] - g[1]), l23 = __floats2bfloat162_rn(e[2] - g[2], e[3] - g[3]);
        // BF16 row layout: per 32-element k-chunk [hi 32 | lo 32]  (kernels_tc.cuh Mode::pe_off)
        __nv_bfloat16 *row = reinterpret_cast<__nv_bfloat16 *>(A.pe) + (size_t)v * (2 * CCSP_H) + (col >> 5) * 64 + (col & 31);
        uint2 hv, lv;
        hv.x = *reinterpret_cast<uint32_t *>(&h01); hv.y = *reinterpret_cast<uint32_t *>(&h23);
        lv.x = *reinterpret_cast<uint32_t *>(&l01); lv.y = *reinterpret_cast<uint32_t *>(&l23);
        *reinterpret_cast<uint2 *>(row) = hv;
        *reinterpret_cast<uint2 *>(row + 32) = lv;
      }
    }
  }
}

// =================================================================================================
// FP32 tiled GEMM core (64 x 128 output tile, BK = 16, 256 threads, 4 x 8 per thread).
//   A rows come from up to three 256-wide segments, each either gathered through an index array
//   (per-edge endpoint embeddings, denoise_fn.py:326-327, 337) or dense.
//   Bt is the transposed weight [K][ldb] (k-major) so tile loads are coalesced.
// =================================================================================================
struct RowSrc {
  const float *src[3];
  const int *idx[3];     // nullptr => dense rows of 256 floats
  int nseg;
};

#define SG_BM 64
#define SG_BN 128
#define SG_BK 16
#define SG_LDA (SG_BM + 4)

__device__ __forceinline__ void simt_gemm_tile(const RowSrc &rs, int m0, const float *__restrict__ Bt, int ldb,
                                               int n0, float (*As)[SG_LDA], float (*Bs)[SG_BN],
                                               float acc[4][8]) {
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int arow = tid & 63, akq = tid >> 6;       // A load: one float4 per thread
  const int bk = tid >> 5, bc4 = tid & 31;         // B load: two float4 per thread
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const int K = rs.nseg * CCSP_H;
  const float *rp = nullptr;
  for (int k0 = 0; k0 < K; k0 += SG_BK) {
    if ((k0 & (CCSP_H - 1)) == 0) {
      int seg = k0 >> 8;
      size_t row = rs.idx[seg] ? (size_t)__ldg(&rs.idx[seg][m0 + arow]) : (size_t)(m0 + arow);
      rp = rs.src[seg] + row * CCSP_H;
    }
    float4 av = *reinterpret_cast<const float4 *>(rp + (k0 & (CCSP_H - 1)) + akq * 4);
    float4 bv0 = __ldg(reinterpret_cast<const float4 *>(Bt + (size_t)(k0 + bk) * ldb + n0 + bc4 * 4));
    float4 bv1 = __ldg(reinterpret_cast<const float4 *>(Bt + (size_t)(k0 + bk + 8) * ldb + n0 + bc4 * 4));
    __syncthreads();
    As[akq * 4 + 0][arow] = av.x;
    As[akq * 4 + 1][arow] = av.y;
    As[akq * 4 + 2][arow] = av.z;
    As[akq * 4 + 3][arow] = av.w;
    *reinterpret_cast<float4 *>(&Bs[bk][bc4 * 4]) = bv0;
    *reinterpret_cast<float4 *>(&Bs[bk + 8][bc4 * 4]) = bv1;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float4 a4 = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
      float4 b0 = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4 *>(&Bs[kk][64 + tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
}

// First layer of mlps[c] over edge tiles (denoise_fn.py:346-356), split by linearity of the
// concatenation:
//   EPI_STATIC : S[e,:]  = W_c[:, static cols] @ [ (grasp_i) ; geom_i ; geom_j ] + b_c         (plan build)
//   EPI_L1     : H[e,:]  = SiLU( W_c[:, pose cols] @ [pose_i ; pose_j] + S[e,:] + tb[t,c,:] )  (every call)
// S lives in the blocked layout of common.cuh::blk_off (shared with the tensor-core epilogue); H of this
// FP32 path is plain row-major.
// grid (E'/64, 4); tile_type[row/128] selects the weight block.
enum { EPI_STATIC = 0, EPI_L1 = 1 };

template <int EPI>
__global__ void __launch_bounds__(256)
k_edge_l1_simt(RowSrc rs, const float *__restrict__ Wt /*[C][K][512]*/, const int *__restrict__ tile_type,
               const float *__restrict__ bias /*[C][512]*/, const float *__restrict__ S /*[E',512]*/,
               const float *__restrict__ tb /*[C][512] at t*/, float *__restrict__ out /*[E',512]*/) {
  __shared__ __align__(16) float As[SG_BK][SG_LDA];
  __shared__ __align__(16) float Bs[SG_BK][SG_BN];
  const int m0 = blockIdx.x * SG_BM, n0 = blockIdx.y * SG_BN;
  const int c = tile_type[m0 / CCSP_TILE_M];
  const int K = rs.nseg * CCSP_H;
  float acc[4][8];
  simt_gemm_tile(rs, m0, Wt + (size_t)c * K * CCSP_H2, CCSP_H2, n0, As, Bs, acc);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const size_t row = (size_t)(m0 + ty * 4 + i);
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      const int col = n0 + hlf * 64 + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = acc[i][hlf * 4 + j];
      if (EPI == EPI_STATIC) {
        float4 bb = __ldg(reinterpret_cast<const float4 *>(bias + (size_t)c * CCSP_H2 + col));
        v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w;
        *reinterpret_cast<float4 *>(out + blk_off(row, col)) = make_float4(v[0], v[1], v[2], v[3]);
        continue;
      } else {
        float4 s4 = __ldg(reinterpret_cast<const float4 *>(S + blk_off(row, col)));
        float4 t4 = __ldg(reinterpret_cast<const float4 *>(tb + (size_t)c * CCSP_H2 + col));
        v[0] = silu_f(v[0] + s4.x + t4.x);
        v[1] = silu_f(v[1] + s4.y + t4.y);
        v[2] = silu_f(v[2] + s4.z + t4.z);
        v[3] = silu_f(v[3] + s4.w + t4.w);
      }
      *reinterpret_cast<float4 *>(out + row * CCSP_H2 + col) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

// Shared pose decoder on both halves of H (denoise_fn.py:253-257, 357-362):
//   rows q = 2 e + slot; A[q,:] = H[e, slot*256 : slot*256+256]  (H viewed as [2E', 256])
//   D1 = SiLU(A @ Wd1^T + bd1) [128];  o[q,:] = D1 @ Wd2^T + bd2 [P]
// grid (2E'/64).
__global__ void __launch_bounds__(256)
k_edge_dec_simt(const float *__restrict__ Hbuf, const float *__restrict__ Wd1t /*[256][128]*/,
                const float *__restrict__ bd1, const float *__restrict__ Wd2 /*[P][128]*/,
                const float *__restrict__ bd2, int P, float *__restrict__ o /*[2E', P]*/) {
  __shared__ __align__(16) float As[SG_BK][SG_LDA];
  __shared__ __align__(16) float Bs[SG_BK][SG_BN];
  __shared__ float D1[SG_BM][SG_BN + 1];
  const int m0 = blockIdx.x * SG_BM;
  RowSrc rs;
  rs.src[0] = Hbuf; rs.idx[0] = nullptr; rs.nseg = 1;
  rs.src[1] = rs.src[2] = nullptr; rs.idx[1] = rs.idx[2] = nullptr;
  float acc[4][8];
  simt_gemm_tile(rs, m0, Wd1t, CCSP_HH, 0, As, Bs, acc);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int col = (j >> 2) * 64 + tx * 4 + (j & 3);
      D1[ty * 4 + i][col] = silu_f(acc[i][j] + __ldg(&bd1[col]));
    }
  __syncthreads();
  for (int idx = threadIdx.x; idx < SG_BM * P; idx += 256) {
    int r = idx / P, p = idx % P;
    float a = 0.f;
    for (int k = 0; k < CCSP_HH; ++k) a = fmaf(D1[r][k], __ldg(&Wd2[p * CCSP_HH + k]), a);
    o[(size_t)(m0 + r) * P + p] = a + __ldg(&bd2[p]);
  }
}

}  // namespace ccsp
