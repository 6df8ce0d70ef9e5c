// kernels_node_tc.cuh — node kernel for the BF16 operand modes with the pose encoder's second layer on tcgen05.
//
// Same contract as k_node (kernels_simt.cuh): deterministic scatter-reduce of the per-(edge, endpoint) decoder
// outputs in the reference's order, 1/sqrt(deg), masked rows, DDPM posterior step / ULA step / init, optional pin,
// history, then pe = pose_encoder(x') written split (hi | lo BF16) for the edge kernel's gather.
//
// Why: in k_node the 128 -> 256 layer is 2.1 M FP32 FMAs per 64 nodes, i.e. >= 16 k cycles per SM on a one-wave
// grid, and the kernel sits between every two edge kernels (22 % of the step in profiles/bench_r1_f).  The kernel is
// latency-bound, so it keeps one wave of small CTAs: a CTA owns 64 nodes (the upper 64 rows of its M = 128 MMA idle).
//   phase 1a 8 threads per node fetch the node's incident decoder outputs in parallel into shared memory (the
//            dependent chain node_ptr -> node_src -> o is paid once, not once per batch of 4);
//   phase 1b one thread per node sums them in the reference's order, normalises, applies the DDPM / ULA / init update
//            (one Philox call per node) and pins;
//   phase 1c every thread evaluates 16 of the 128 first-layer outputs of its row and stores them as two 16-byte
//            pieces of the A operand (hi and lo BF16, SWIZZLE_64B) in shared memory;
//   phase 2  one thread issues the 128 x 256 x 128 GEMM (3-term split, FP32 accumulate in TMEM) against W2, which
//            a single thread fetched with cp.async.bulk while phase 1 ran;
//   phase 3  8 warps: TMEM -> + b2 -> SiLU -> hi/lo -> the node's row of pe_split.
#pragma once
#include "kernels_fused2.cuh"
#include "kernels_simt.cuh"

namespace ccsp {
namespace tc {

template <class M>
struct NodeTcCfg {
  static_assert(M::KIND == KIND_BF16, "BF16 operand modes only");
  static constexpr int ROWS = 64;                             // nodes per CTA (rows 64..127 of the M = 128 MMA are idle)
  static constexpr int NKC = CCSP_HH / M::KC;                 // 128 / 32 = 4 k-chunks
  static constexpr int B_STAGE = M::NS * CCSP_H * ROWB;       // 256 weight rows x 64 B (x2 parts): 32 KB
  static constexpr int OFF_B = NKC * M::A_STAGE;
  static constexpr int OFF_EXTRA = OFF_B + NKC * B_STAGE;
  static constexpr int ROW_THREADS = 512, THREADS = ROW_THREADS + 32;
  static constexpr int PARTS = ROW_THREADS / ROWS;            // 8 threads per node
  static constexpr int STAGE_ENTRIES = 32;                    // incident (edge, endpoint) rows staged per node (rest: fallback)
  static_assert(ROWS * STAGE_ENTRIES * 32 <= NKC * M::A_STAGE || M::NS == 1, "scratch must fit in the A operand region");
  // barriers 64 | xs [64][8] | w0 [128][8] | b0 [128] | b2 [256]
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
  // scratch for the scatter-reduce: [64 rows][32 entries][8 floats], aliases the (not yet written) A operand region;
  // the NS == 1 layout has only 32 KB there, so it stages 4 floats per entry (P <= 4) or falls back
  float *scratch = reinterpret_cast<float *>(smem);
  constexpr int EW = M::NS == 2 ? 8 : 4;                      // floats per staged entry

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * C::ROWS;
  const int P = A.P;
  const uint32_t smem_base = smem_u32(smem);
  long long *const tr = (A.trace && blockIdx.x == 1 && (tid == 0 || tid == C::ROW_THREADS)) ? A.trace + (tid == 0 ? 0 : 16) : nullptr;
#define NTR(slot) do { if (tr) tr[slot] = clock64(); } while (0)
  NTR(0);
  if (tid == 0) pdl_launch_dependents();

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
  pdl_wait();                            // o / x / pe are produced (or still read) by the preceding edge kernel

  const int r = tid & (C::ROWS - 1), part = tid >> 6;         // row threads: node row, 1 of 8 helpers of that row
  const int v = row0 + r;
  const bool reduce = tid < C::ROW_THREADS && v < A.n && A.mode != NODE_INIT && A.mode != NODE_ENCODE;
  int k0 = 0, k1 = 0;
  bool masked = false;
  float x_old[CCSP_MAXP], z_in[CCSP_MAXP], aux[CCSP_MAXP], gtv[CCSP_MAXP];   // loaded early by the node's summing thread (part == 0)
#pragma unroll
  for (int p = 0; p < CCSP_MAXP; ++p) { x_old[p] = 0.f; z_in[p] = 0.f; aux[p] = 0.f; gtv[p] = 0.f; }
  // ---- phase 1a: the 8 threads of a node fetch its incident decoder outputs in parallel -------------------------
  if (tid < C::ROW_THREADS) {
    if (tid < C::ROWS && v < A.n && A.mode != NODE_ENCODE) {
      const size_t ix = (size_t)v * P;
      masked = A.mask[v] != 0;
#pragma unroll
      for (int p = 0; p < CCSP_MAXP; ++p) {
        if (p >= P) continue;
        if (A.mode != NODE_INIT) x_old[p] = A.x[ix + p];
        if (A.z) z_in[p] = A.z[ix + p];
        if (masked) aux[p] = (A.mode != NODE_INIT) ? A.xtail[ix + p] : 0.f;   // eps of a masked row = x[:, -P:]
        if (masked && A.pin) gtv[p] = A.gt[ix + p];
      }
    }
    if (reduce) {
      masked = A.mask[v] != 0;
      if (!masked && P <= EW) {
        k0 = A.node_ptr[v]; k1 = A.node_ptr[v + 1];
        const int kend = min(k1, k0 + C::STAGE_ENTRIES);
        int src[4];
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = k0 + part + j * C::PARTS;
          if (k < kend) { src[j] = A.node_src[k]; cnt = j + 1; }
        }
        float *dst = scratch + ((size_t)r * C::STAGE_ENTRIES + part) * EW;
        if (P == 4) {
          float4 a[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) if (j < cnt) a[j] = reinterpret_cast<const float4 *>(A.o)[src[j]];
#pragma unroll
          for (int j = 0; j < 4; ++j) if (j < cnt) *reinterpret_cast<float4 *>(dst + j * C::PARTS * EW) = a[j];
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (j < cnt) {
            const float *orow = A.o + (size_t)src[j] * P;
#pragma unroll
            for (int p = 0; p < EW; ++p) if (p < P) dst[j * C::PARTS * EW + p] = orow[p];
          }
        }
      }
    }
  }
  NTR(1);
  tc_fence_before();
  __syncthreads();                       // staged rows, w0s/b0s/b2s, barrier init, TMEM address
  tc_fence_after();
  NTR(2);
  const uint32_t tmem_base = *tmem_ptr;

  // ---- phase 1b: one thread per node: ordered sum, normalise, update, pin ------------------------------------------
  if (tid < C::ROWS) {
    float xn[CCSP_MAXP];
#pragma unroll
    for (int p = 0; p < CCSP_MAXP; ++p) xn[p] = 0.f;
    if (v < A.n) {
      const size_t ix = (size_t)v * P;
      if (A.mode == NODE_ENCODE) {
        _Pragma("unroll") for (int p = 0; p < CCSP_MAXP; ++p) if (p < P) xn[p] = A.x_in[ix + p];
      } else {
        float zz[CCSP_MAXP];
#pragma unroll
        for (int p = 0; p < CCSP_MAXP; ++p) zz[p] = z_in[p];
        const bool need_z = (A.mode == NODE_DDPM || A.mode == NODE_ULA || (A.mode == NODE_INIT && !A.has_xinit));
        if (need_z && !A.z) philox_normals(A.seed, A.draw, A.node_offset + (unsigned long long)v, P, zz);
        float eps[CCSP_MAXP];
#pragma unroll
        for (int p = 0; p < CCSP_MAXP; ++p) eps[p] = 0.f;
        if (A.mode != NODE_INIT) {
          if (masked) {
            _Pragma("unroll") for (int p = 0; p < CCSP_MAXP; ++p) if (p < P) eps[p] = aux[p];
          } else {
            // same accumulation order as the reference's scatter_add_ (sequential adds per component)
            int k = k0;
            if (P <= EW) {
              const int kend = min(k1, k0 + C::STAGE_ENTRIES);
              const float *srow = scratch + (size_t)r * C::STAGE_ENTRIES * EW;
              for (; k < kend; ++k, srow += EW) {
                if (EW == 8) {
                  const float4 a0 = *reinterpret_cast<const float4 *>(srow), a1 = *reinterpret_cast<const float4 *>(srow + 4);
                  eps[0] = __fadd_rn(eps[0], a0.x); eps[1] = __fadd_rn(eps[1], a0.y); eps[2] = __fadd_rn(eps[2], a0.z); eps[3] = __fadd_rn(eps[3], a0.w);
                  if (P > 4) { eps[4] = __fadd_rn(eps[4], a1.x); eps[5] = __fadd_rn(eps[5], a1.y); eps[6] = __fadd_rn(eps[6], a1.z); eps[7] = __fadd_rn(eps[7], a1.w); }
                } else {
                  const float4 a0 = *reinterpret_cast<const float4 *>(srow);
                  eps[0] = __fadd_rn(eps[0], a0.x); eps[1] = __fadd_rn(eps[1], a0.y); eps[2] = __fadd_rn(eps[2], a0.z); eps[3] = __fadd_rn(eps[3], a0.w);
                }
              }
            } else {
              k0 = A.node_ptr[v]; k1 = A.node_ptr[v + 1]; k = k0;
            }
            for (; k < k1; ++k) {            // entries beyond the staged window (high-degree nodes)
              const float *orow = A.o + (size_t)A.node_src[k] * P;
              _Pragma("unroll") for (int p = 0; p < CCSP_MAXP; ++p) if (p < P) eps[p] = __fadd_rn(eps[p], orow[p]);
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
          const float xv = x_old[p];
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
          if (A.pin && masked) xo = gtv[p];
          A.x[ix + p] = xo;
          if (A.hist) A.hist[ix + p] = xo;
          xn[p] = xo;
        }
      }
    }
#pragma unroll
    for (int p = 0; p < CCSP_MAXP; ++p) xs[r][p] = xn[p];
  }
  NTR(3);
  __syncthreads();                       // xs; the scratch (= A operand region) is free again
  NTR(4);

  if (tid < C::ROW_THREADS) {
    // ---- phase 1c: 16 first-layer outputs of row r -> two 16-byte pieces of k-chunk part/2 of the A operand ------
    float xr[CCSP_MAXP];
#pragma unroll
    for (int d = 0; d < CCSP_MAXP; ++d) xr[d] = xs[r][d];
    uint8_t *a_stage = smem + (part >> 1) * M::A_STAGE;
#pragma unroll
    for (int qq = 0; qq < 2; ++qq) {
      const int q = (part & 1) * 2 + qq;
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = (part >> 1) * 32 + q * 8 + i;
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
  NTR(5);
  __syncthreads();                       // A operand complete (rows 64..127 are never read back)
  NTR(6);

  if (tid == C::ROW_THREADS) {
    // ---- phase 2: D[128 x 256] = h . W2^T ------------------------------------------------------------------
    mbar_wait(bfull, 0);
    NTR(7);
    tc_fence_after();
#pragma unroll
    for (int kc = 0; kc < C::NKC; ++kc) {
      const uint32_t a_hi = smem_base + kc * M::A_STAGE;
      issue_chunk<M, CCSP_H>(tmem_base, a_hi, smem_base + C::OFF_B + kc * C::B_STAGE, kc == 0);
    }
    umma_commit(tfull);
    NTR(8);
  } else if (tid < C::ROW_THREADS && (warp & 3) < 2) {
    // ---- phase 3: warp w <-> TMEM lanes 32 (w & 3).. (only lanes 0..63 hold nodes), columns 64 (w >> 2).. +63 ----
    const int quarter = warp & 3, cg = warp >> 2;
    const int rr = quarter * 32 + lane, vv = row0 + rr;
    mbar_wait(tfull, 0);
    NTR(7);
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
  NTR(8 + (tid == 0 ? 0 : 1));
  tc_fence_before();
  __syncthreads();
  NTR(10);
  if (warp == C::ROW_THREADS / 32) tmem_dealloc(tmem_base, 256);
  NTR(11);
#undef NTR
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
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(C::THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // see pdl_wait() in kernels_fused2.cuh
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k_node_tc<M>, a, w2_blob);
}

}  // namespace tc
}  // namespace ccsp
