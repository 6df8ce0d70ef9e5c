// kernels_node_tc.cuh — node kernel for the BF16 operand modes with the pose encoder's second layer on tcgen05.
//
// Same contract as k_node (kernels_simt.cuh): deterministic scatter-reduce of the per-(edge, endpoint) decoder
// outputs in the reference's order, 1/sqrt(deg), masked rows, DDPM posterior step / ULA step / init, optional pin,
// history, then pe = pose_encoder(x') written split (hi | lo BF16) for the edge kernel's gather.
//
// Why: in k_node the 128 -> 256 layer is 2.1 M FP32 FMAs per 64 nodes, i.e. >= 16 k cycles per SM on a one-wave
// grid, and the kernel sits between every two edge kernels (22 % of the step in profiles/bench_r1_f).  Here a CTA owns
// 128 nodes:
//   phase 1  thread (row r, quarter kq): kq == 0 threads reduce + update their node (all P components; one Philox
//            call per node), then every thread evaluates 32 of the 128 first-layer outputs of its row and stores
//            them as one 64-byte row of k-chunk kq of the A operand (hi and lo BF16, SWIZZLE_64B) in shared memory;
//   phase 2  one thread issues the 128 x 256 x 128 GEMM (3-term split, FP32 accumulate in TMEM) against W2, which
//            a single thread fetched with cp.async.bulk while phase 1 ran;
//   phase 3  16 warps: TMEM -> + b2 -> SiLU -> hi/lo -> the node's row of pe_split.
#pragma once
#include "kernels_fused2.cuh"
#include "kernels_simt.cuh"

namespace ccsp {
namespace tc {

template <class M>
struct NodeTcCfg {
  static_assert(M::KIND == KIND_BF16, "BF16 operand modes only");
  static constexpr int ROWS = SUB_M;                          // nodes per CTA
  static constexpr int NKC = CCSP_HH / M::KC;                 // 128 / 32 = 4 k-chunks
  static constexpr int B_STAGE = M::NS * CCSP_H * ROWB;       // 256 weight rows x 64 B (x2 parts): 32 KB
  static constexpr int OFF_B = NKC * M::A_STAGE;
  static constexpr int OFF_EXTRA = OFF_B + NKC * B_STAGE;
  static constexpr int ROW_THREADS = 512, THREADS = ROW_THREADS + 32;
  // barriers 64 | xs [128][8] | w0 [128][8] | b0 [128] | b2 [256]
  static constexpr int SMEM_EXTRA = 64 + (ROWS * CCSP_MAXP + CCSP_HH * CCSP_MAXP + CCSP_HH + CCSP_H) * 4;
  static constexpr int SMEM_BYTES = OFF_EXTRA + SMEM_EXTRA + 1024;
  static_assert(SMEM_BYTES <= 227 * 1024, "node kernel does not fit in shared memory");
};

template <class M>
__global__ void __launch_bounds__(NodeTcCfg<M>::THREADS, 1) k_node_tc(const NodeArgs A, const uint8_t *__restrict__ w2_blob) {
  using C = NodeTcCfg<M>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *extra = smem + C::OFF_EXTRA;
  uint64_t *bfull = reinterpret_cast<uint64_t *>(extra);      // W2 landed
  uint64_t *tfull = bfull + 1;                                // accumulator complete
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tfull + 1);
  float (*xs)[CCSP_MAXP] = reinterpret_cast<float (*)[CCSP_MAXP]>(extra + 64);
  float *w0s = reinterpret_cast<float *>(extra + 64) + C::ROWS * CCSP_MAXP;   // [128][8]
  float *b0s = w0s + CCSP_HH * CCSP_MAXP;                                      // [128]
  float *b2s = b0s + CCSP_HH;                                                  // [256]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * C::ROWS;
  const int P = A.P;
  const uint32_t smem_base = smem_u32(smem);

  if (tid == C::ROW_THREADS) {
    mbar_init(bfull, 1); mbar_init(tfull, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(bfull, C::NKC * C::B_STAGE);
#pragma unroll
    for (int kc = 0; kc < C::NKC; ++kc)
      bulk_g2s(smem_base + C::OFF_B + kc * C::B_STAGE, w2_blob + (size_t)kc * C::B_STAGE, C::B_STAGE, bfull);
  }
  if (warp == C::ROW_THREADS / 32) tmem_alloc(tmem_ptr, 256);
  if (tid < C::ROW_THREADS) {
    for (int i = tid; i < CCSP_HH * CCSP_MAXP; i += C::ROW_THREADS) {
      const int j = i / CCSP_MAXP, d = i % CCSP_MAXP;
      w0s[i] = d < P ? __ldg(&A.W0[j * P + d]) : 0.f;
    }
    if (tid < CCSP_HH) b0s[tid] = __ldg(&A.b0[tid]);
    if (tid < CCSP_H) b2s[tid] = __ldg(&A.b2[tid]);
  }

  const int r = tid & (C::ROWS - 1), kq = tid >> 7;           // row threads: node row, first-layer output quarter
  const int v = row0 + r;
  // ---- phase 1a: reduce + update, one node per thread (all P components) ---------------------------------
  if (tid < C::ROWS) {
    float xn[CCSP_MAXP];
#pragma unroll
    for (int p = 0; p < CCSP_MAXP; ++p) xn[p] = 0.f;
    if (v < A.n) {
      const size_t ix = (size_t)v * P;
      const bool masked = A.mask[v] != 0;
      if (A.mode == NODE_ENCODE) {
        _Pragma("unroll") for (int p = 0; p < CCSP_MAXP; ++p) if (p < P) xn[p] = A.x_in[ix + p];
      } else {
        float zz[CCSP_MAXP];
#pragma unroll
        for (int p = 0; p < CCSP_MAXP; ++p) zz[p] = 0.f;
        const bool need_z = (A.mode == NODE_DDPM || A.mode == NODE_ULA || (A.mode == NODE_INIT && !A.has_xinit));
        if (need_z) {
          if (A.z) { _Pragma("unroll") for (int p = 0; p < CCSP_MAXP; ++p) if (p < P) zz[p] = A.z[ix + p]; }
          else philox_normals(A.seed, A.draw, A.node_offset + (unsigned long long)v, P, zz);
        }
        float eps[CCSP_MAXP];
#pragma unroll
        for (int p = 0; p < CCSP_MAXP; ++p) eps[p] = 0.f;
        if (A.mode != NODE_INIT) {
          if (masked) {
            _Pragma("unroll") for (int p = 0; p < CCSP_MAXP; ++p) if (p < P) eps[p] = A.xtail[ix + p];
          } else {
            // same accumulation order as the reference's scatter_add_ (sequential adds per component)
            const int k0 = A.node_ptr[v], k1 = A.node_ptr[v + 1];
            if (P == 4) {
              const float4 *o4 = reinterpret_cast<const float4 *>(A.o);
              float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
              int k = k0;
              for (; k + 4 <= k1; k += 4) {
                const int s0 = A.node_src[k], s1 = A.node_src[k + 1], s2 = A.node_src[k + 2], s3 = A.node_src[k + 3];
                const float4 a0 = o4[s0], a1 = o4[s1], a2 = o4[s2], a3 = o4[s3];
                acc.x = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc.x, a0.x), a1.x), a2.x), a3.x);
                acc.y = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc.y, a0.y), a1.y), a2.y), a3.y);
                acc.z = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc.z, a0.z), a1.z), a2.z), a3.z);
                acc.w = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(acc.w, a0.w), a1.w), a2.w), a3.w);
              }
              for (; k < k1; ++k) {
                const float4 a0 = o4[A.node_src[k]];
                acc.x = __fadd_rn(acc.x, a0.x); acc.y = __fadd_rn(acc.y, a0.y);
                acc.z = __fadd_rn(acc.z, a0.z); acc.w = __fadd_rn(acc.w, a0.w);
              }
              eps[0] = acc.x; eps[1] = acc.y; eps[2] = acc.z; eps[3] = acc.w;
            } else {
              for (int k = k0; k < k1; ++k) {
                const float *orow = A.o + (size_t)A.node_src[k] * P;
                _Pragma("unroll") for (int p = 0; p < CCSP_MAXP; ++p) if (p < P) eps[p] = __fadd_rn(eps[p], orow[p]);
              }
            }
            if (A.normalize) {
              const float sd = sqrtf((float)(k1 - k0));
              _Pragma("unroll") for (int p = 0; p < CCSP_MAXP; ++p) if (p < P) eps[p] = eps[p] / sd;
            }
          }
        }
#pragma unroll
        for (int p = 0; p < CCSP_MAXP; ++p) {
          if (p >= P) continue;
          const float xv = (A.mode == NODE_INIT) ? 0.f : A.x[ix + p];
          float xo;
          if (A.mode == NODE_INIT) {
            xo = A.has_xinit ? A.x_in[ix + p] : __fmul_rn(0.5f, zz[p]);
          } else if (A.mode == NODE_DDPM) {
            const float x0 = __fsub_rn(__fmul_rn(A.a, xv), __fmul_rn(A.b, eps[p]));
            const float mean = __fadd_rn(__fmul_rn(A.c1, x0), __fmul_rn(A.c2, xv));
            xo = __fadd_rn(mean, __fmul_rn(A.sigma, zz[p]));
          } else {  // NODE_ULA
            const float grad = __fmul_rn(-eps[p], A.gscale);
            xo = __fadd_rn(__fadd_rn(xv, __fmul_rn(grad, A.ss)), __fmul_rn(zz[p], A.std));
          }
          if (A.pin && masked) xo = A.gt[ix + p];
          A.x[ix + p] = xo;
          if (A.hist) A.hist[ix + p] = xo;
          xn[p] = xo;
        }
      }
    }
#pragma unroll
    for (int p = 0; p < CCSP_MAXP; ++p) xs[r][p] = xn[p];
  }
  tc_fence_before();
  __syncthreads();                       // xs, w0s/b0s/b2s, barrier init, TMEM address
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (tid < C::ROW_THREADS) {
    // ---- phase 1b: 32 first-layer outputs of row r -> the 64-byte row of k-chunk kq of the A operand ---------
    float xr[CCSP_MAXP];
#pragma unroll
    for (int d = 0; d < CCSP_MAXP; ++d) xr[d] = xs[r][d];
    uint8_t *a_stage = smem + kq * M::A_STAGE;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = kq * 32 + q * 8 + i;
        const float4 wa = *reinterpret_cast<const float4 *>(&w0s[j * CCSP_MAXP]);
        const float4 wb = *reinterpret_cast<const float4 *>(&w0s[j * CCSP_MAXP + 4]);
        float acc = 0.f;                 // fmaf chain in the order d = 0..P-1 (the zero-padded tail adds exact zeros)
        acc = fmaf(xr[0], wa.x, acc); acc = fmaf(xr[1], wa.y, acc); acc = fmaf(xr[2], wa.z, acc); acc = fmaf(xr[3], wa.w, acc);
        if (P > 4) { acc = fmaf(xr[4], wb.x, acc); acc = fmaf(xr[5], wb.y, acc); acc = fmaf(xr[6], wb.z, acc); acc = fmaf(xr[7], wb.w, acc); }
        f[i] = silu_raw(acc + b0s[j]);
      }
      uint4 hi, lo;
      split_pair(f[0], f[1], hi.x, lo.x); split_pair(f[2], f[3], hi.y, lo.y);
      split_pair(f[4], f[5], hi.z, lo.z); split_pair(f[6], f[7], hi.w, lo.w);
      uint8_t *dst = a_stage + sw64_off(r, q);
      *reinterpret_cast<uint4 *>(dst) = hi;
      if (M::NS == 2) *reinterpret_cast<uint4 *>(dst + PART) = lo;
    }
    fence_proxy_async();
  }
  __syncthreads();                       // A operand complete

  if (tid == C::ROW_THREADS) {
    // ---- phase 2: D[128 x 256] = h . W2^T ------------------------------------------------------------------
    mbar_wait(bfull, 0);
    tc_fence_after();
#pragma unroll
    for (int kc = 0; kc < C::NKC; ++kc) {
      const uint32_t a_hi = smem_base + kc * M::A_STAGE;
      issue_chunk<M, CCSP_H>(tmem_base, a_hi, smem_base + C::OFF_B + kc * C::B_STAGE, kc == 0);
    }
    umma_commit(tfull);
  } else if (tid < C::ROW_THREADS) {
    // ---- phase 3: warp w <-> TMEM lanes 32 (w & 3).., columns 64 (w >> 2).. +63 -----------------------------
    const int quarter = warp & 3, cg = warp >> 2;
    const int rr = quarter * 32 + lane, vv = row0 + rr;
    mbar_wait(tfull, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + cg * 64 + ((uint32_t)(quarter * 32) << 16);
    uint8_t *prow = reinterpret_cast<uint8_t *>(A.pe) + (size_t)vv * M::PE_ROW_BYTES;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float vals[32];
      tmem_ld32(taddr + half * 32, vals);
      const int c0 = cg * 64 + half * 32;
      uint4 hi[4], lo[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          f[i] = silu_raw(vals[q * 8 + i] + b2s[c0 + q * 8 + i]);
          if (vv == A.n) f[i] = 0.f;     // zero row read by padded edges
        }
        split_pair(f[0], f[1], hi[q].x, lo[q].x); split_pair(f[2], f[3], hi[q].y, lo[q].y);
        split_pair(f[4], f[5], hi[q].z, lo[q].z); split_pair(f[6], f[7], hi[q].w, lo[q].w);
      }
      if (vv <= A.n) {
        uint8_t *dh = prow + c0 * 2;
        stg256(dh, hi[0], hi[1]); stg256(dh + 32, hi[2], hi[3]);
        stg256(dh + M::PE_LO_OFF, lo[0], lo[1]); stg256(dh + M::PE_LO_OFF + 32, lo[2], lo[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == C::ROW_THREADS / 32) tmem_dealloc(tmem_base, 256);
}

template <class M>
cudaError_t launch_node_tc(const NodeArgs &a, const uint8_t *w2_blob, cudaStream_t st) {
  using C = NodeTcCfg<M>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_node_tc<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const unsigned blocks = (unsigned)((a.n + 1 + C::ROWS - 1) / C::ROWS);
  k_node_tc<M><<<blocks, C::THREADS, C::SMEM_BYTES, st>>>(a, w2_blob);
  return cudaGetLastError();
}

}  // namespace tc
}  // namespace ccsp
