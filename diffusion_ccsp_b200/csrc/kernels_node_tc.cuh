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
//   phase 3  16 warps: TMEM -> + b2 -> SiLU -> hi/lo -> the node's row of pe_split.
#pragma once
#include "kernels_fused2.cuh"
#include "kernels_simt.cuh"

namespace ccsp {
namespace tc {

template <class M>
struct NodeTcCfg {
  static_assert(M::KIND == KIND_BF16, "BF16 operand modes only");
  static constexpr int ROWS = 64;                             // nodes per CTA (rows 64..127 of the M = 128 MMA repeat them)
  static constexpr int NKC = CCSP_HH / M::KC;                 // 128 / 32 = 4 k-chunks
  static constexpr int B_STAGE = M::NS * CCSP_H * ROWB;       // 256 weight rows x 64 B (x2 parts): 32 KB
  static constexpr int OFF_B = NKC * M::A_STAGE;
  static constexpr int OFF_EXTRA = OFF_B + NKC * B_STAGE;
  static constexpr int ROW_THREADS = 512, THREADS = ROW_THREADS + 32;
  static constexpr int PARTS = ROW_THREADS / ROWS;            // 8 threads per node
  static constexpr int STAGE_ENTRIES = 31;                    // incident (edge, endpoint) rows staged per node (rest: fallback)
  static constexpr int EW = M::NS == 2 ? 8 : 4;               // floats per staged entry (the NS == 1 A region is half as big)
  // floats per node row of the scratch; = 28 (mod 32), so the 32 lanes (= 32 nodes) of a warp hit distinct banks
  static constexpr int ROW_STRIDE = STAGE_ENTRIES * EW + (EW == 8 ? 4 : 0);
  static_assert(ROW_STRIDE % 32 == 28 && ROWS * ROW_STRIDE * 4 <= NKC * M::A_STAGE, "scratch layout");
  // barriers 64 | xs [64][8] | zs [64][8] | w0 [128][8] | b0 [128] | b2 [256]
  static constexpr int SMEM_EXTRA = 64 + (2 * ROWS * CCSP_MAXP + CCSP_HH * CCSP_MAXP + CCSP_HH + CCSP_H) * 4;
  static constexpr int SMEM_BYTES = OFF_EXTRA + SMEM_EXTRA + 1024;
  static_assert(SMEM_BYTES <= 227 * 1024, "node kernel does not fit in shared memory");
};

// PT: pose width known at compile time (2 boxes, 4 qualitative / triangles, 5 robot; 0 = read A.P at run time).  The
// kernel runs once per launch from a cold instruction cache with two warps on its critical path, so the code it
// has to fetch is its run time: a compile-time P removes every `p < P` predicate and the dead component code.
// PERSIST = true: all A.num_iters iterations of a sample() in one launch (see NodeArgs): barriers, TMEM, W2 and the first-layer
// tables are set up once; the kernel boundary is replaced by the edge_done / node_done flags.
template <class M, int PT, bool PERSIST = false>
__global__ void __launch_bounds__(NodeTcCfg<M>::THREADS, 1) k_node_tc(const NodeArgs A0, const uint8_t *__restrict__ w2_blob) {
  using C = NodeTcCfg<M>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *extra = smem + C::OFF_EXTRA;
  uint64_t *bfull = reinterpret_cast<uint64_t *>(extra);      // W2 landed
  uint64_t *tfull = bfull + 1;                                // accumulator complete
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tfull + 1);
  float (*xs)[CCSP_MAXP] = reinterpret_cast<float (*)[CCSP_MAXP]>(extra + 64);
  float (*zs)[CCSP_MAXP] = reinterpret_cast<float (*)[CCSP_MAXP]>(extra + 64 + C::ROWS * CCSP_MAXP * 4);   // noise per node
  float *w0s = reinterpret_cast<float *>(extra + 64) + 2 * C::ROWS * CCSP_MAXP;   // [128][8]
  float *b0s = w0s + CCSP_HH * CCSP_MAXP;                                      // [128]
  float *b2s = b0s + CCSP_HH;                                                  // [256]
  // scratch for the scatter-reduce: [64 rows][32 entries][8 floats], aliases the (not yet written) A operand region;
  // the NS == 1 layout has only 32 KB there, so it stages 4 floats per entry (P <= 4) or falls back
  float *scratch = reinterpret_cast<float *>(smem);
  constexpr int EW = C::EW;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int P = PT ? PT : A0.P;
  const uint32_t smem_base = smem_u32(smem);
  long long *const tr = (!PERSIST && A0.trace && blockIdx.x == 1 && (tid == 0 || tid == C::ROW_THREADS)) ? A0.trace + (tid == 0 ? 0 : 16) : nullptr;
  unsigned long long *const ptr_ = (PERSIST && blockIdx.x == 0 && tid == 0) ? reinterpret_cast<unsigned long long *>(A0.trace) : nullptr;
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
      w0s[i] = d < P ? __ldg(&A0.W0[j * P + d]) : 0.f;
    }
    if (tid < CCSP_HH) b0s[tid] = __ldg(&A0.b0[tid]);
    if (tid < CCSP_H) b2s[tid] = __ldg(&A0.b2[tid]);
  }
  const int r = tid & (C::ROWS - 1), part = tid >> 6;         // row threads: node row, 1 of 8 helpers of that row
  const int n_iters = PERSIST ? A0.num_iters : 1;
  // persistent mode: every iteration visits the plan's chains in turn (NodeArgs::num_chains), and within a chain this CTA
  // takes the 64-row blocks blockIdx.x, + gridDim.x, ..  (launch-per-iteration mode: one chain, one block per CTA)
  const int n_ch = (PERSIST && A0.num_chains > 1) ? A0.num_chains : 1;
  uint32_t nblk_done = 0;                // blocks this CTA has pushed through the tensor core (phase of tfull)
#pragma unroll 1
  for (int iter = 0; iter < n_iters; ++iter) {
#pragma unroll 1
  for (int ch = 0; ch < n_ch; ++ch) {
  const bool part_mode = PERSIST && n_ch > 1 && A0.node_partition;
  if (part_mode && ch != (int)(blockIdx.x % n_ch)) continue;               // this CTA serves one chain only
  const int blk0 = part_mode ? (int)(blockIdx.x / n_ch) : (int)blockIdx.x;
  const int blk_step = part_mode ? (int)(gridDim.x / n_ch) : (int)gridDim.x;
  NodeArgs A = A0;                       // this iteration's arguments
  const int crow0 = n_ch > 1 ? A0.chain_row0[ch] : 0;                     // rows of this chain: [crow0, cend)
  const int cend = n_ch > 1 ? A0.chain_row0[ch + 1] : A0.n;
  const int ccopy_end = n_ch > 1 ? cend : A0.n + 1;    // (+ the zero row n; the persistent edge kernel zero-fills padded rows itself)
  const int nblk = (ccopy_end - crow0 + C::ROWS - 1) / C::ROWS;
  bool need_wait = PERSIST && iter > 0;  // the edge kernel's evaluation iter - 1 of this chain has to be complete
  const int tq = iter * n_ch + ch;       // trace slot
  PTRACE(ptr_, 0, tq);
  if (PERSIST) {
    const NodeEval E = A0.sched[iter];
    A.mode = E.mode; A.pin = E.pin; A.a = E.a; A.b = E.b; A.c1 = E.c1; A.c2 = E.c2; A.sigma = E.sigma;
    A.gscale = E.gscale; A.ss = E.ss; A.std = E.std; A.draw = E.draw;
    A.z = A0.z ? A0.z + (size_t)E.draw * A0.nP : nullptr;
    A.hist = (A0.hist && E.hist_slot >= 0) ? A0.hist + (size_t)E.hist_slot * A0.nP : nullptr;
  }
#pragma unroll 1
  for (int blk = blk0; blk < nblk; blk += PERSIST ? blk_step : nblk) {
  const int row0 = crow0 + blk * C::ROWS;
  const int v = row0 + r;
  if (tid < C::ROW_THREADS) {
    if (part == 1 && v < cend && !A.z &&
        (A.mode == NODE_DDPM || A.mode == NODE_ULA || (A.mode == NODE_INIT && !A.has_xinit))) {
      float zz[CCSP_MAXP];               // the node's Philox draw, computed off the summing thread's critical path
#pragma unroll
      for (int p = 0; p < CCSP_MAXP; ++p) zz[p] = 0.f;
      philox_normals(A.seed, A.draw, A.node_offset + (unsigned long long)v, P, zz);
#pragma unroll
      for (int p = 0; p < CCSP_MAXP; ++p) zs[r][p] = zz[p];
    }
  }
  // Everything below this line and above pdl_wait() reads only plan constants (graph structure, mask, pinned poses,
  // injected noise), so it runs under the tail of the preceding edge kernel: after the wait one dependent load level
  // (the decoder outputs o, and x) is left on the critical path instead of three.
  const bool reduce = tid < C::ROW_THREADS && v < cend && A.mode != NODE_INIT && A.mode != NODE_ENCODE;
  int k0 = 0, k1 = 0;
  bool masked = false;
  float x_old[CCSP_MAXP], z_in[CCSP_MAXP], aux[CCSP_MAXP], gtv[CCSP_MAXP];   // state of the node's summing thread (part == 0)
#pragma unroll
  for (int p = 0; p < CCSP_MAXP; ++p) { x_old[p] = 0.f; z_in[p] = 0.f; aux[p] = 0.f; gtv[p] = 0.f; }
  int src[4] = {0, 0, 0, 0};
  int cnt = 0;
  if (tid < C::ROW_THREADS && v < cend && A.mode != NODE_ENCODE) {
    masked = A.mask[v] != 0;
    if (tid < C::ROWS) {
      const size_t ix = (size_t)v * P;
#pragma unroll
      for (int p = 0; p < CCSP_MAXP; ++p) {
        if (p >= P) continue;
        if (A.z) z_in[p] = A.z[ix + p];
        if (masked) aux[p] = (A.mode != NODE_INIT) ? A.xtail[ix + p] : 0.f;   // eps of a masked row = x[:, -P:]
        if (masked && A.pin) gtv[p] = A.gt[ix + p];
      }
    }
    if (reduce && !masked && P <= EW) {
      k0 = A.node_ptr[v]; k1 = A.node_ptr[v + 1];
      const int kend = min(k1, k0 + C::STAGE_ENTRIES);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + part + j * C::PARTS;
        if (k < kend) { src[j] = A.node_src[k]; cnt = j + 1; }
      }
    }
  }
  if (!PERSIST) {
    pdl_wait();                          // o / x / pe are produced (or still read) by the preceding edge kernel
  } else if (need_wait) {                // ... or by evaluation iter - 1 of the persistent edge kernel
    if (tid == 0) wait_flag_ge(A0.edge_done + 32 * ch, (unsigned)iter * (n_ch > 1 ? 2u * A0.chain_units[ch] : A0.edge_ctas));
    __syncthreads();
    need_wait = false;
    PTRACE(ptr_, 1, tq);
  }

  // ---- phase 1a: the 8 threads of a node fetch its incident decoder outputs in parallel -------------------------
  if (tid < C::ROW_THREADS) {
    if (tid < C::ROWS && v < cend && A.mode != NODE_ENCODE && A.mode != NODE_INIT) {
      const size_t ix = (size_t)v * P;
#pragma unroll
      for (int p = 0; p < CCSP_MAXP; ++p)
        if (p < P) x_old[p] = A.x[ix + p];
    }
    if (cnt > 0) {
      float *dst = scratch + (size_t)r * C::ROW_STRIDE + part * EW;
      if (P == 4) {
        float4 a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) if (j < cnt) a[j] = __ldcg(reinterpret_cast<const float4 *>(A.o) + src[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) if (j < cnt) *reinterpret_cast<float4 *>(dst + j * C::PARTS * EW) = a[j];
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (j < cnt) {
          const float *orow = A.o + (size_t)src[j] * P;
#pragma unroll
          for (int p = 0; p < EW; ++p) if (p < P) dst[j * C::PARTS * EW + p] = __ldcg(orow + p);
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
    if (v < cend) {
      const size_t ix = (size_t)v * P;
      if (A.mode == NODE_ENCODE) {
        _Pragma("unroll") for (int p = 0; p < CCSP_MAXP; ++p) if (p < P) xn[p] = A.x_in[ix + p];
      } else {
        float zz[CCSP_MAXP];
#pragma unroll
        for (int p = 0; p < CCSP_MAXP; ++p) zz[p] = z_in[p];
        const bool need_z = (A.mode == NODE_DDPM || A.mode == NODE_ULA || (A.mode == NODE_INIT && !A.has_xinit));
        if (need_z && !A.z) {
#pragma unroll
          for (int p = 0; p < CCSP_MAXP; ++p) zz[p] = zs[r][p];
        }
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
              const float *srow = scratch + (size_t)r * C::ROW_STRIDE;
              if (P <= 4) {                // loads batched by 4, adds strictly in order
                for (; k + 4 <= kend; k += 4, srow += 4 * EW) {
                  const float4 a0 = *reinterpret_cast<const float4 *>(srow), a1 = *reinterpret_cast<const float4 *>(srow + EW);
                  const float4 a2 = *reinterpret_cast<const float4 *>(srow + 2 * EW), a3 = *reinterpret_cast<const float4 *>(srow + 3 * EW);
                  eps[0] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(eps[0], a0.x), a1.x), a2.x), a3.x);
                  eps[1] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(eps[1], a0.y), a1.y), a2.y), a3.y);
                  eps[2] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(eps[2], a0.z), a1.z), a2.z), a3.z);
                  eps[3] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(eps[3], a0.w), a1.w), a2.w), a3.w);
                }
              }
#pragma unroll 1
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
#pragma unroll 1
            for (; k < k1; ++k) {            // entries beyond the staged window (high-degree nodes)
              const float *orow = A.o + (size_t)A.node_src[k] * P;
              _Pragma("unroll") for (int p = 0; p < CCSP_MAXP; ++p) if (p < P) eps[p] = __fadd_rn(eps[p], __ldcg(orow + p));
            }
            NTR(12);
            if (A.normalize) {
              const float sd = sqrtf((float)(k1 - k0));
              _Pragma("unroll") for (int p = 0; p < CCSP_MAXP; ++p) if (p < P) eps[p] = eps[p] / sd;
            }
          }
        }
        NTR(13);
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
    NTR(14);
#pragma unroll
    for (int p = 0; p < CCSP_MAXP; ++p) xs[r][p] = xn[p];
  }
  NTR(3);
  PTRACE(ptr_, 2, tq);
  __syncthreads();                       // xs; the scratch (= A operand region) is free again
  NTR(4);

  if (tid < C::ROW_THREADS) {
    // ---- phase 1c: 16 first-layer outputs of row r -> two 16-byte pieces of k-chunk part/2 of the A operand ------
    float xr[CCSP_MAXP];
#pragma unroll
    for (int d = 0; d < CCSP_MAXP; ++d) xr[d] = xs[r][d];
    uint8_t *a_stage = smem + (part >> 1) * M::A_STAGE;
#pragma unroll 1                          // (rolled: this kernel runs from a cold instruction cache, code size is time)
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
      // rows 64..127 of the M = 128 MMA repeat the 64 nodes: TMEM lanes 64..127 then hold the same results, and the warps that
      // own those lanes take half of every node's columns in phase 3 (a warp can only read its own 32 lanes)
      uint8_t *dup = a_stage + sw64_off(r + C::ROWS, q);
      *reinterpret_cast<uint4 *>(dup) = hi;
      if (M::NS == 2) *reinterpret_cast<uint4 *>(dup + PART) = lo;
    }
    fence_proxy_async();
  }
  NTR(5);
  __syncthreads();                       // A operand complete
  NTR(6);

  if (tid == C::ROW_THREADS) {
    // ---- phase 2: D[128 x 256] = h . W2^T ------------------------------------------------------------------
    mbar_wait(bfull, 0);               // (later iterations: phase 0 stays complete)
    NTR(7);
    tc_fence_after();
#pragma unroll
    for (int kc = 0; kc < C::NKC; ++kc) {
      const uint32_t a_hi = smem_base + kc * M::A_STAGE;
      issue_chunk<M, CCSP_H>(tmem_base, a_hi, smem_base + C::OFF_B + kc * C::B_STAGE, kc == 0);
    }
    umma_commit(tfull);
    NTR(8);
  } else if (tid < C::ROW_THREADS) {
    // ---- phase 3: warp w <-> TMEM lanes 32 (w & 3).. = node (w & 1) * 32 + lane (lanes 64..127 repeat the nodes), columns
    //      64 (w >> 2) + 32 ((w & 3) >> 1) .. +31: all 16 warps work, 32 columns each ----
    const int quarter = warp & 3, cg = warp >> 2, chalf = quarter >> 1;
    const int rr = (quarter & 1) * 32 + lane, vv = row0 + rr;
    mbar_wait(tfull, nblk_done & 1);
    NTR(7);
    tc_fence_after();
    const uint32_t taddr = tmem_base + cg * 64 + chalf * 32 + ((uint32_t)(quarter * 32) << 16);
    // results are staged row-major in shared memory (the A operand region is free now) and written out coalesced:
    // a lane-per-row store of 16 B at a 1 KB stride costs one L1 pass per lane, 4096 of them per CTA
    uint8_t *orow = smem + (size_t)rr * M::PE_ROW_BYTES;
#pragma unroll 1                          // rolled, 16 columns per iteration (see phase 1c)
    for (int c16 = 0; c16 < 2; ++c16) {
      float vals[16];
      tmem_ld16(taddr + c16 * 16, vals);
#pragma unroll
      for (int h8 = 0; h8 < 2; ++h8) {
        const int c0 = cg * 64 + chalf * 32 + c16 * 16 + h8 * 8;
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          f[i] = silu_raw(vals[h8 * 8 + i] + b2s[c0 + i]);
          if (vv == A.n) f[i] = 0.f;     // zero row read by padded edges
        }
        uint4 hi, lo;
        split_pair(f[0], f[1], hi.x, lo.x); split_pair(f[2], f[3], hi.y, lo.y);
        split_pair(f[4], f[5], hi.z, lo.z); split_pair(f[6], f[7], hi.w, lo.w);
        // 16-byte chunk c of the row lives at chunk (c + row) & 63: the 32 lanes (= 32 rows) hit distinct banks
        const int ch = c0 >> 3;          // hi chunk index 0..31; lo = +32
        *reinterpret_cast<uint4 *>(orow + (((ch + rr) & 63) << 4)) = hi;
        *reinterpret_cast<uint4 *>(orow + (((ch + 32 + rr) & 63) << 4)) = lo;
      }
    }
  }
  NTR(8 + (tid == 0 ? 0 : 1));
  tc_fence_before();
  __syncthreads();
  if (tid < C::ROW_THREADS) {
    // ---- phase 4: 16 warps copy the 64 rows of pe_split out, one coalesced 512-byte store per warp and half row ----
#pragma unroll 1
    for (int i = 0; i < C::ROWS / 16; ++i) {
      const int rr = warp * (C::ROWS / 16) + i, vv = row0 + rr;
      if (vv >= ccopy_end) break;
      const uint8_t *src = smem + (size_t)rr * M::PE_ROW_BYTES;
      uint8_t *dst = reinterpret_cast<uint8_t *>(A.pe) + (size_t)vv * M::PE_ROW_BYTES;
      // global 16-byte slot g of the row holds piece (g & 3) of part ((g >> 2) & 1) of k-chunk (g >> 3)  (Mode::pe_off);
      // the staged row is [hi chunks 0..31 | lo chunks 32..63], rotated by the row index
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int gs = lane + 32 * h;
        const int c = ((gs >> 2) & 1) * 32 + (gs >> 3) * 4 + (gs & 3);
        *reinterpret_cast<uint4 *>(dst + (gs << 4)) = *reinterpret_cast<const uint4 *>(src + (((c + rr) & 63) << 4));
      }
    }
  }
  NTR(10);
  ++nblk_done;
  if (PERSIST) __syncthreads();          // the staging rows (= the next block's scratch) are free again
  }
  if (PERSIST && n_ch == 1) {            // x / history / pe of this iteration are written: tell the edge kernel
    if (need_wait) {                     // (a CTA without a block still keeps step with the edge kernel: the flag counts CTAs)
      if (tid == 0) wait_flag_ge(A0.edge_done, (unsigned)iter * A0.edge_ctas);
      __syncthreads();
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) red_release_gpu_add(A0.node_done, 1u);
    PTRACE(ptr_, 3, tq);
  } else if (PERSIST && blk0 < nblk) {   // chains: the flag counts blocks; a CTA without a block of this chain stays out of it
    __threadfence();
    __syncthreads();
    if (tid == 0) red_release_gpu_add(A0.node_done + 32 * ch, (unsigned)((nblk - blk0 + blk_step - 1) / blk_step));
    PTRACE(ptr_, 3, tq);
  }
  }
  }
  if (warp == C::ROW_THREADS / 32) tmem_dealloc(*tmem_ptr, 256);
  NTR(11);
#undef NTR
}

template <class M, int PT>
cudaError_t launch_node_tc_p(const NodeArgs &a, const uint8_t *w2_blob, cudaStream_t st) {
  using C = NodeTcCfg<M>;
  static bool configured_dev[64] = {};      // per device: function attributes belong to the device's context
  int dev_ = 0;
  cudaGetDevice(&dev_);
  bool &configured = configured_dev[dev_ & 63];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_node_tc<M, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
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
  return cudaLaunchKernelEx(&cfg, k_node_tc<M, PT>, a, w2_blob);
}

// Persistent variant (run-time pose width: the code stays hot in the instruction cache anyway), no PDL attribute.
// (configure = load the function NOW: with lazy module loading the first launch of a kernel can synchronise the device, which
// would deadlock against the already running persistent edge kernel that is waiting for this one)
template <class M>
cudaError_t configure_node_tc_persistent() {
  using C = NodeTcCfg<M>;
  static bool configured_dev[64] = {};
  int dev_ = 0;
  cudaGetDevice(&dev_);
  if (!configured_dev[dev_ & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k_node_tc<M, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, k_node_tc<M, 0, true>);
    if (e != cudaSuccess) return e;
    configured_dev[dev_ & 63] = true;
  }
  return cudaSuccess;
}
template <class M>
cudaError_t launch_node_tc_persistent(const NodeArgs &a, const uint8_t *w2_blob, cudaStream_t st, int ctas = 0) {
  using C = NodeTcCfg<M>;
  cudaError_t e0 = configure_node_tc_persistent<M>();
  if (e0 != cudaSuccess) return e0;
  if (a.num_iters <= 0 || !a.sched) return cudaErrorInvalidValue;
  const unsigned blocks = ctas > 0 ? (unsigned)ctas : (unsigned)((a.n + 1 + C::ROWS - 1) / C::ROWS);
  k_node_tc<M, 0, true><<<blocks, C::THREADS, C::SMEM_BYTES, st>>>(a, w2_blob);
  return cudaGetLastError();
}

template <class M>
cudaError_t launch_node_tc(const NodeArgs &a, const uint8_t *w2_blob, cudaStream_t st) {
  switch (a.P) {
    case 2: return launch_node_tc_p<M, 2>(a, w2_blob, st);
    case 4: return launch_node_tc_p<M, 4>(a, w2_blob, st);
    case 5: return launch_node_tc_p<M, 5>(a, w2_blob, st);
    default: return launch_node_tc_p<M, 0>(a, w2_blob, st);
  }
}

}  // namespace tc
}  // namespace ccsp
