// kernels_tc.cuh — tcgen05 (5th-gen tensor core) kernels for the two dense per-edge layers.
//
// One persistent, warp-specialised GEMM kernel template, C[256 x NTILE] tiles with FP32 accumulators in
// TMEM, operands staged in shared memory in the canonical K-major SWIZZLE_64B layout:
//
//   warps 0-7   epilogue   tcgen05.ld accumulator -> registers -> fused epilogue -> global
//   warps 8-11  A producer gather pose-embedding rows through the edge index (16 B loads from L2),
//                          split FP32 -> (hi, lo) operand pair, st.shared into the swizzled layout
//   warp  12    B loader   weights are pre-split / pre-swizzled per k-chunk on the host, one
//                          cp.async.bulk (TMA engine, no tensor map needed) per stage
//   warp  13    MMA issuer tcgen05.mma (one elected lane), tcgen05.commit -> mbarriers; owns TMEM alloc
//
// FP32 fidelity: the reference computes these layers in true FP32 (cuBLAS/oneDNN sgemm).  The tensor
// cores take TF32/BF16 operands, so every FP32 operand x is split as x = hi + lo (+ residual) and
//   A.B ~= A_lo.B_hi + A_hi.B_lo + A_hi.B_hi            (3 MMAs, FP32 accumulation in TMEM)
// which recovers ~22 (TF32x3) / ~16 (BF16x3) mantissa bits per operand; single-pass modes are kept as
// explicitly lower-precision options (include/ccsp_b200.h CcspMath).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace ccsp {
namespace tc {

enum { KIND_TF32 = 0, KIND_BF16 = 1 };
enum { EPI_TC_L1 = 0, EPI_TC_DEC = 1 };

constexpr int ROWB = 64;        // bytes of K per operand row per stage (SWIZZLE_64B span)
constexpr int SUB_M = 128;      // rows per tcgen05.mma (one TMEM accumulator = 128 lanes)
constexpr int NUM_EPI_WARPS = 8, NUM_PROD_WARPS = 4;
constexpr int WARP_PROD0 = NUM_EPI_WARPS, WARP_BLOAD = WARP_PROD0 + NUM_PROD_WARPS, WARP_MMA = WARP_BLOAD + 1;
constexpr int NUM_THREADS = (WARP_MMA + 1) * 32;   // 448
constexpr uint32_t SPIN_LIMIT = 1u << 27;   // bounded spin: a protocol bug traps instead of hanging the GPU

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

template <int KIND>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (KIND == KIND_TF32) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void sts128(uint32_t saddr, const uint4 &v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 ldg_nc_f4(const float4 *p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
// SiLU with the hardware exp2/rcp approximations (abs error < 3e-7 |silu(x)| + 1e-7);
// the FP32 validation path keeps the libm-accurate version (common.cuh silu_f).
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// 32 lanes x 32 consecutive FP32 columns: thread i of the warp gets lane (lane_base + i), columns col..col+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_64B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (=1, unused for swizzled K-major) | [32,46) SBO >> 4
//   (8 rows x 64 B = 512 B) | [46,48) version = 1 (Blackwell) | [61,64) layout type = 4 (SWIZZLE_64B)
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// byte offset of 16-byte chunk `q` (0..3) of row `r` inside a SWIZZLE_64B operand tile (Swizzle<2,4,3>)
__host__ __device__ __forceinline__ uint32_t sw64_off(uint32_t r, uint32_t q) { return r * ROWB + ((q ^ ((r >> 1) & 3)) << 4); }

// FP32 -> operand-pair splits
__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&v);
}

// ---------------------------------------------------------------------------------------------------
// Tile shape: one CTA tile = 256 operand rows (two 128-row sub-tiles = two TMEM accumulators) x NTILE
// output columns.  Both sub-tiles share every B (weight) stage, which halves the L2->SM weight traffic
// per FLOP relative to a 128-row tile — the resource that bounds this kernel.
// ---------------------------------------------------------------------------------------------------
template <int KIND_, int NSPLIT_, int NTILE_, int EPI_>
struct Cfg {
  static constexpr int KIND = KIND_, NSPLIT = NSPLIT_, NTILE = NTILE_, EPI = EPI_;
  static constexpr int ELT = KIND == KIND_TF32 ? 4 : 2;      // operand element bytes
  static constexpr int KC = ROWB / ELT;                      // K elements per stage (16 / 32)
  static constexpr int UMMA_K = 32 / ELT;                    // K per tcgen05.mma (8 / 16)
  static constexpr int KSTEPS = KC / UMMA_K;                 // 2
  static constexpr int NS = NSPLIT == 3 ? 2 : 1;             // operand copies per stage (hi[, lo])
  static constexpr int SUB = 2;                              // 128-row sub-tiles per CTA tile
  static constexpr int TILE_ROWS = SUB * SUB_M;              // 256
  static constexpr int A_PART = SUB_M * ROWB;                // 8 KB
  static constexpr int A_SUB = NS * A_PART;
  static constexpr int A_STAGE = SUB * A_SUB;
  static constexpr int B_PART = NTILE * ROWB;
  static constexpr int B_STAGE = NS * B_PART;
  static constexpr int STAGE = A_STAGE + B_STAGE;
  static constexpr int NSTAGE = STAGE <= 48 * 1024 ? 4 : 3;  // smem ring depth
  static constexpr int ACC_COLS = SUB * NTILE;               // TMEM columns per tile
  static constexpr int NBUF = 512 / ACC_COLS;                // accumulator buffers (1 or 2)
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_EXTRA = 8192;                    // barriers, tmem ptr, epilogue constants
  static constexpr int SMEM_BYTES = NSTAGE * STAGE + SMEM_EXTRA + 1024;   // + alignment slack
  static constexpr uint32_t IDESC = (1u << 4) | ((KIND == KIND_TF32 ? 2u : 1u) << 7) | ((KIND == KIND_TF32 ? 2u : 1u) << 10) |
                                    ((uint32_t)(NTILE >> 3) << 17) | ((uint32_t)(SUB_M >> 4) << 24);
  static_assert(NBUF >= 1 && SMEM_BYTES <= 227 * 1024, "tile does not fit");
};

struct GemmArgs {
  // A operand: rows of `nseg` 256-float segments, each gathered through an index array or dense
  const float *a_src[2];
  const int *a_idx[2];
  int nseg;
  // B operand: host-packed stage blobs [group][n_tile][k-chunk][B_STAGE bytes]
  const uint8_t *b_blob;
  const int *tile_type;      // weight group per 256-row tile (nullptr -> group 0)
  int num_m_tiles, n_tiles;  // 256-row tiles, NTILE-column tiles
  // EPI_TC_L1: H[row, nt*NTILE + j] = SiLU(acc + S[row, .] + tb[group, .])
  const float *S, *tb;
  float *H;
  // EPI_TC_DEC: d = SiLU(acc + bd1); o[row, p] = sum_j d_j Wd2[p, j] + bd2[p]
  const float *bd1, *Wd2, *bd2;
  int P;
  float *o;
  // performance ablations for the harness (results are garbage when non-zero):
  //   1 = producers skip the gather/convert/store, 2 = B loader skips the bulk copy,
  //   4 = epilogue skips global loads/stores, 8 = MMA warp skips the tcgen05.mma instructions
  int dbg;
};

template <class C>
__global__ void __launch_bounds__(NUM_THREADS, 1) k_gemm_tc(const GemmArgs A) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *extra = smem + C::NSTAGE * C::STAGE;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(extra);            // [NSTAGE]
  uint64_t *empty_bar = full_bar + 4;                                  // [NSTAGE]
  uint64_t *tfull_bar = empty_bar + 4;                                 // [NBUF]
  uint64_t *tempty_bar = tfull_bar + 2;                                // [NBUF]
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tempty_bar + 2);
  float *epi_const = reinterpret_cast<float *>(extra + 256);           // L1: tb slice [NTILE]; DEC: bd1[128] + w2t[128*8] + bd2[8]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NKC = A.nseg * CCSP_H / C::KC;
  const int num_tiles = A.num_m_tiles * A.n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::NSTAGE; ++s) { mbar_init(&full_bar[s], 32 + 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < C::NBUF; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], NUM_EPI_WARPS * 32); }
    fence_barrier_init();
  }
  if (warp == WARP_MMA) tmem_alloc(tmem_ptr, C::TMEM_COLS);
  if (C::EPI == EPI_TC_DEC && warp < NUM_EPI_WARPS) {
    const int nth = NUM_EPI_WARPS * 32;
    for (int i = threadIdx.x; i < CCSP_HH; i += nth) epi_const[i] = A.bd1[i];
    for (int i = threadIdx.x; i < CCSP_MAXP * CCSP_HH; i += nth) {      // w2t[j][p], zero-padded to 8 outputs
      const int j = i / CCSP_MAXP, pp = i % CCSP_MAXP;
      epi_const[CCSP_HH + i] = pp < A.P ? A.Wd2[pp * CCSP_HH + j] : 0.f;
    }
    for (int i = threadIdx.x; i < CCSP_MAXP; i += nth) epi_const[CCSP_HH + CCSP_MAXP * CCSP_HH + i] = i < A.P ? A.bd2[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp >= WARP_PROD0 && warp < WARP_PROD0 + C::NSTAGE) {
    // ================================ A producers ================================================
    // producer warp pw OWNS ring stage pw and converts k-chunks kc == pw (mod NSTAGE): 256 rows x 64 B of
    // operand per chunk.  One owner per stage keeps every waiter at most one mbarrier phase ahead (parity
    // waits cannot tell phases two apart); with a 3-deep ring the 4th producer warp stays idle.
    const int pw = warp - WARP_PROD0;
    const uint32_t smem_base = smem_u32(smem);
    const int q = lane & 3, r0 = lane >> 2;
    uint32_t tile_iter = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
      const int m0 = (tile / A.n_tiles) * C::TILE_ROWS;
      int cur_seg = -1;
      uint32_t rowoff[32];
      const int kc_first = (pw + C::NSTAGE - (int)((tile_iter * (uint32_t)NKC) % C::NSTAGE)) % C::NSTAGE;
      for (int kc = kc_first; kc < NKC; kc += C::NSTAGE) {
        const uint32_t g = tile_iter * NKC + kc;             // global chunk counter; g % NSTAGE == pw
        const uint32_t s = pw;
        const int k0 = kc * C::KC;
        const int seg = k0 >> 8;
        if (seg != cur_seg) {
          cur_seg = seg;
#pragma unroll
          for (int p = 0; p < 32; ++p) {
            const int row = m0 + p * 8 + r0;
            rowoff[p] = (uint32_t)(A.a_idx[seg] ? __ldg(&A.a_idx[seg][row]) : row) * CCSP_H;
          }
        }
        const float *src = A.a_src[seg] + (k0 & (CCSP_H - 1));
        const uint32_t stA = smem_base + s * C::STAGE;
        mbar_wait(&empty_bar[s], ((g / C::NSTAGE) & 1) ^ 1);
        if (A.dbg & 1) {
          mbar_arrive(&full_bar[s]);
          continue;
        }
        if (C::KIND == KIND_TF32) {
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {                 // two batches of 16 rows-in-flight per lane
            float4 v[16];
#pragma unroll
            for (int p = 0; p < 16; ++p) v[p] = __ldg(reinterpret_cast<const float4 *>(src + rowoff[hb * 16 + p] + q * 4));
#pragma unroll
            for (int p = 0; p < 16; ++p) {
              const int r = (hb * 16 + p) * 8 + r0;          // row within the 256-row tile
              const uint32_t dst = stA + (r >> 7) * C::A_SUB + sw64_off(r & 127, q);
              uint4 hi;
              hi.x = tf32_rna(v[p].x); hi.y = tf32_rna(v[p].y); hi.z = tf32_rna(v[p].z); hi.w = tf32_rna(v[p].w);
              sts128(dst, hi);
              if (C::NS == 2) {
                uint4 lo;
                lo.x = tf32_rna(v[p].x - __uint_as_float(hi.x)); lo.y = tf32_rna(v[p].y - __uint_as_float(hi.y));
                lo.z = tf32_rna(v[p].z - __uint_as_float(hi.z)); lo.w = tf32_rna(v[p].w - __uint_as_float(hi.w));
                sts128(dst + C::A_PART, lo);
              }
            }
          }
        } else {
#pragma unroll
          for (int hb = 0; hb < 4; ++hb) {                 // four batches of 8 rows (2 float4 each) per lane
            float4 v[16];
#pragma unroll
            for (int p = 0; p < 8; ++p) {
              const float *sp = src + rowoff[hb * 8 + p] + q * 8;
              v[2 * p] = __ldg(reinterpret_cast<const float4 *>(sp));
              v[2 * p + 1] = __ldg(reinterpret_cast<const float4 *>(sp + 4));
            }
#pragma unroll
            for (int p = 0; p < 8; ++p) {
              const int r = (hb * 8 + p) * 8 + r0;
              const uint32_t dst = stA + (r >> 7) * C::A_SUB + sw64_off(r & 127, q);
              const float f[8] = {v[2 * p].x, v[2 * p].y, v[2 * p].z, v[2 * p].w, v[2 * p + 1].x, v[2 * p + 1].y, v[2 * p + 1].z, v[2 * p + 1].w};
              float h[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) h[i] = __bfloat162float(__float2bfloat16_rn(f[i]));
              uint4 hi;
              hi.x = pack_bf16(h[0], h[1]); hi.y = pack_bf16(h[2], h[3]); hi.z = pack_bf16(h[4], h[5]); hi.w = pack_bf16(h[6], h[7]);
              sts128(dst, hi);
              if (C::NS == 2) {
                uint4 lo;
                lo.x = pack_bf16(f[0] - h[0], f[1] - h[1]); lo.y = pack_bf16(f[2] - h[2], f[3] - h[3]);
                lo.z = pack_bf16(f[4] - h[4], f[5] - h[5]); lo.w = pack_bf16(f[6] - h[6], f[7] - h[7]);
                sts128(dst + C::A_PART, lo);
              }
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(&full_bar[s]);
      }
    }
  } else if (warp == WARP_BLOAD) {
    // ================================ B loader ==================================================
    if (lane == 0) {
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / A.n_tiles, nt = tile % A.n_tiles;
        const int grp = A.tile_type ? __ldg(&A.tile_type[mt]) : 0;
        const uint8_t *blob = A.b_blob + ((size_t)(grp * A.n_tiles + nt) * NKC) * C::B_STAGE;
        for (int kc = 0; kc < NKC; ++kc, ++g) {
          const uint32_t s = g % C::NSTAGE;
          mbar_wait(&empty_bar[s], ((g / C::NSTAGE) & 1) ^ 1);
          if (A.dbg & 2) {
            mbar_arrive(&full_bar[s]);
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[s], C::B_STAGE);
          bulk_g2s(smem + s * C::STAGE + C::A_STAGE, blob + (size_t)kc * C::B_STAGE, C::B_STAGE, &full_bar[s]);
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ================================ MMA issuer ================================================
    if (lane == 0) {
      uint32_t g = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t buf = tcount % C::NBUF;
        mbar_wait(&tempty_bar[buf], ((tcount / C::NBUF) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * C::ACC_COLS;
        for (int kc = 0; kc < NKC; ++kc, ++g) {
          const uint32_t s = g % C::NSTAGE;
          mbar_wait(&full_bar[s], (g / C::NSTAGE) & 1);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + s * C::STAGE);
          const uint32_t b_hi = a0 + C::A_STAGE, b_lo = b_hi + C::B_PART;
#pragma unroll
          for (int ks = 0; ks < ((A.dbg & 8) ? 0 : C::KSTEPS); ++ks) {
            const uint32_t ko = ks * 32;
#pragma unroll
            for (int sub = 0; sub < C::SUB; ++sub) {
              const uint32_t a_hi = a0 + sub * C::A_SUB, a_lo = a_hi + C::A_PART;
              const uint32_t d = d_tmem + sub * C::NTILE;
              if (C::NSPLIT == 3) {
                umma<C::KIND>(d, smem_desc_sw64(a_lo + ko), smem_desc_sw64(b_hi + ko), C::IDESC, (kc | ks) != 0);
                umma<C::KIND>(d, smem_desc_sw64(a_hi + ko), smem_desc_sw64(b_lo + ko), C::IDESC, 1);
                umma<C::KIND>(d, smem_desc_sw64(a_hi + ko), smem_desc_sw64(b_hi + ko), C::IDESC, 1);
              } else {
                umma<C::KIND>(d, smem_desc_sw64(a_hi + ko), smem_desc_sw64(b_hi + ko), C::IDESC, (kc | ks) != 0);
              }
            }
          }
          umma_commit(&empty_bar[s]);        // frees the smem stage once these MMAs have read it
        }
        umma_commit(&tfull_bar[buf]);        // accumulators complete -> epilogue
      }
    }
  } else if (warp < NUM_EPI_WARPS) {
    // ============== epilogue: warp w <-> sub-tile (w >> 2), TMEM lanes 32 (w & 3) .. +31 ==========
    uint32_t tcount = 0;
    const int quarter = warp & 3, sub = warp >> 2;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t buf = tcount % C::NBUF;
      const int mt = tile / A.n_tiles, nt = tile % A.n_tiles;
      const size_t row = (size_t)mt * C::TILE_ROWS + sub * SUB_M + quarter * 32 + lane;
      const uint32_t taddr = tmem_base + buf * C::ACC_COLS + sub * C::NTILE + ((uint32_t)(quarter * 32) << 16);
      if (C::EPI == EPI_TC_L1) {
        const int grp = A.tile_type ? __ldg(&A.tile_type[mt]) : 0;
        const float4 *Sv = reinterpret_cast<const float4 *>(A.S + row * CCSP_H2 + nt * C::NTILE);
        float4 *Hv = reinterpret_cast<float4 *>(A.H + row * CCSP_H2 + nt * C::NTILE);
        float4 sn[8];                                       // static term, prefetched one 32-column chunk ahead
#pragma unroll
        for (int qq = 0; qq < 8; ++qq) sn[qq] = (A.dbg & 4) ? make_float4(0.f, 0.f, 0.f, 0.f) : ldg_nc_f4(Sv + qq);
        asm volatile("bar.sync 1, 256;" ::: "memory");      // previous tile's readers of epi_const are done
        for (int i = threadIdx.x; i < C::NTILE; i += NUM_EPI_WARPS * 32)
          epi_const[i] = __ldg(&A.tb[(size_t)grp * CCSP_H2 + nt * C::NTILE + i]);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        mbar_wait(&tfull_bar[buf], (tcount / C::NBUF) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int cb = 0; cb < C::NTILE; cb += 32) {
          float4 sc[8];
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) sc[qq] = sn[qq];
          if (cb + 32 < C::NTILE && !(A.dbg & 4)) {
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) sn[qq] = ldg_nc_f4(Sv + (cb + 32) / 4 + qq);
          }
          float v[32];
          tmem_ld32(taddr + cb, v);
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) {
            const float4 t4 = *reinterpret_cast<const float4 *>(&epi_const[cb + 4 * qq]);
            float4 o4;
            o4.x = silu_fast(v[4 * qq] + sc[qq].x + t4.x); o4.y = silu_fast(v[4 * qq + 1] + sc[qq].y + t4.y);
            o4.z = silu_fast(v[4 * qq + 2] + sc[qq].z + t4.z); o4.w = silu_fast(v[4 * qq + 3] + sc[qq].w + t4.w);
            if (!(A.dbg & 4) || o4.x == 12345.678f) Hv[cb / 4 + qq] = o4;
          }
        }
      } else {
        mbar_wait(&tfull_bar[buf], (tcount / C::NBUF) & 1);
        tc_fence_after();
        const float *bd1 = epi_const, *w2t = epi_const + CCSP_HH, *bd2 = epi_const + CCSP_HH + CCSP_MAXP * CCSP_HH;
        if (A.P <= 4) {
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
          for (int cb = 0; cb < C::NTILE; cb += 32) {
            float v[32];
            tmem_ld32(taddr + cb, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float d = silu_fast(v[j] + bd1[cb + j]);
              const float4 w = *reinterpret_cast<const float4 *>(&w2t[(cb + j) * CCSP_MAXP]);
              acc.x = fmaf(d, w.x, acc.x); acc.y = fmaf(d, w.y, acc.y); acc.z = fmaf(d, w.z, acc.z); acc.w = fmaf(d, w.w, acc.w);
            }
          }
          const float r4[4] = {acc.x + bd2[0], acc.y + bd2[1], acc.z + bd2[2], acc.w + bd2[3]};
          if (A.P == 4) {
            *reinterpret_cast<float4 *>(A.o + row * 4) = make_float4(r4[0], r4[1], r4[2], r4[3]);
          } else {
#pragma unroll
            for (int p = 0; p < 4; ++p)
              if (p < A.P) A.o[row * A.P + p] = r4[p];
          }
        } else {
          float acc[CCSP_MAXP];
#pragma unroll
          for (int p = 0; p < CCSP_MAXP; ++p) acc[p] = 0.f;
#pragma unroll 1
          for (int cb = 0; cb < C::NTILE; cb += 32) {
            float v[32];
            tmem_ld32(taddr + cb, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float d = silu_fast(v[j] + bd1[cb + j]);
              const float4 w0 = *reinterpret_cast<const float4 *>(&w2t[(cb + j) * CCSP_MAXP]);
              const float4 w1 = *reinterpret_cast<const float4 *>(&w2t[(cb + j) * CCSP_MAXP + 4]);
              acc[0] = fmaf(d, w0.x, acc[0]); acc[1] = fmaf(d, w0.y, acc[1]); acc[2] = fmaf(d, w0.z, acc[2]); acc[3] = fmaf(d, w0.w, acc[3]);
              acc[4] = fmaf(d, w1.x, acc[4]); acc[5] = fmaf(d, w1.y, acc[5]); acc[6] = fmaf(d, w1.z, acc[6]); acc[7] = fmaf(d, w1.w, acc[7]);
            }
          }
#pragma unroll
          for (int p = 0; p < CCSP_MAXP; ++p)
            if (p < A.P) A.o[row * A.P + p] = acc[p] + bd2[p];
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------
// host: pack a row-major FP32 weight block W[n_rows_total, ldw] (nn.Linear layout, K along columns
// starting at k_begin, K_total columns) into stage blobs for the B operand:
//   out[n_tile][kc] = [hi: NTILE rows x 64 B, SWIZZLE_64B][lo: same]     (lo omitted when NSPLIT == 1)
// ---------------------------------------------------------------------------------------------------
inline uint32_t host_tf32_rna(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return u & 0xFFFFE000u;
  u += 0x1000u;
  return u & 0xFFFFE000u;
}
inline uint16_t host_bf16_rn(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
inline float host_bf16_to_f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

template <class C>
void pack_b_blob(const float *W, int ldw, int k_begin, int K_total, int n_rows_total, uint8_t *out) {
  const int n_tiles = n_rows_total / C::NTILE, NKC = K_total / C::KC;
  const int EPC = 16 / C::ELT;   // elements per 16-byte chunk
  for (int nt = 0; nt < n_tiles; ++nt)
    for (int kc = 0; kc < NKC; ++kc) {
      uint8_t *st = out + ((size_t)nt * NKC + kc) * C::B_STAGE;
      for (int r = 0; r < C::NTILE; ++r)
        for (int kk = 0; kk < C::KC; ++kk) {
          const float x = W[(size_t)(nt * C::NTILE + r) * ldw + k_begin + kc * C::KC + kk];
          const uint32_t off = sw64_off(r, kk / EPC) + (kk % EPC) * C::ELT;
          if (C::KIND == KIND_TF32) {
            uint32_t hi = host_tf32_rna(x);
            float hf;
            memcpy(&hf, &hi, 4);
            memcpy(st + off, &hi, 4);
            if (C::NS == 2) { uint32_t lo = host_tf32_rna(x - hf); memcpy(st + C::B_PART + off, &lo, 4); }
          } else {
            uint16_t hi = host_bf16_rn(x);
            memcpy(st + off, &hi, 2);
            if (C::NS == 2) { uint16_t lo = host_bf16_rn(x - host_bf16_to_f(hi)); memcpy(st + C::B_PART + off, &lo, 2); }
          }
        }
    }
}

template <class C>
cudaError_t launch_gemm_tc(const GemmArgs &a, int num_sms, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm_tc<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int tiles = a.num_m_tiles * a.n_tiles;
  if (tiles == 0) return cudaSuccess;
  const int grid = tiles < num_sms ? tiles : num_sms;
  k_gemm_tc<C><<<grid, NUM_THREADS, C::SMEM_BYTES, st>>>(a);
  return cudaGetLastError();
}

}  // namespace tc
}  // namespace ccsp
