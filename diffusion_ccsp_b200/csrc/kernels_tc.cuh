// kernels_tc.cuh — tcgen05 (5th-gen tensor core) kernels for the two dense per-edge layers.
//
//   k_edge_l1_tc   H = SiLU( [pe_i ; pe_j] W_c,pose^T + S + tb[t,c] )      128 edges x 256 outputs per tile, K = 512
//   k_edge_dec_tc  o = SiLU( H_half Wd1^T + bd1 ) Wd2^T + bd2               128 (edge,slot) rows x 128, K = 256
//
// Both are persistent, warp-specialised kernels with FP32 accumulators in TMEM and operands staged in
// shared memory in the canonical K-major SWIZZLE_64B layout.  Every operand reaches shared memory by an
// asynchronous copy of data that is ALREADY in operand format — the conversions happen where the data
// is produced:
//   * pose embeddings are written split (hi | lo) by the node kernel; the first-layer A tile is a
//     16-byte-granular cp.async gather through the edge index straight into the swizzled stage;
//   * the first-layer epilogue writes H split and pre-swizzled per k-chunk, so the decoder's A tile is
//     one contiguous cp.async.bulk (TMA engine) per stage;
//   * weights are split / swizzled per k-chunk on the host once (pack_b_blob), one cp.async.bulk each.
//
//   warps 0-7   epilogue    tcgen05.ld accumulator -> registers -> fused epilogue -> global (coalesced)
//   warps 8-11  A gather    (first layer only) cp.async 16 B pieces, 2 chunks in flight per thread
//   next warp   bulk loader one elected lane: cp.async.bulk of B (and A for the decoder) + expect_tx
//   next warp   MMA issuer  one elected lane: tcgen05.mma / tcgen05.commit -> mbarriers; owns TMEM
//
// FP32 fidelity: the reference computes these layers in true FP32.  Tensor cores take TF32/BF16
// operands, so every FP32 operand x is split as x = hi + lo (+ residual) and
//   A.B ~= A_lo.B_hi + A_hi.B_lo + A_hi.B_hi            (3 MMAs, FP32 accumulation in TMEM);
// single-pass modes are kept as explicitly lower-precision options (include/ccsp_b200.h CcspMath).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace ccsp {
namespace tc {

enum { KIND_TF32 = 0, KIND_BF16 = 1 };

constexpr int ROWB = 64;                 // bytes of K per operand row per stage (SWIZZLE_64B span)
constexpr int SUB_M = 128;               // rows per tile = TMEM lanes of one accumulator
constexpr int PART = SUB_M * ROWB;       // bytes of one operand part (hi or lo) of an A stage: 8 KB
constexpr int NUM_EPI_WARPS = 8;
constexpr int EPI_THREADS = NUM_EPI_WARPS * 32;
// Bounded spin is a DEBUG aid (build with CCSP_DEBUG=1 -> -DCCSP_DEBUG_SPIN_TRAP): a protocol bug then traps instead of
// hanging the GPU.  Release builds spin without a limit: a legitimate stall (debugger, MPS time slice, preemption) must not
// poison the host's CUDA context.
constexpr uint32_t SPIN_LIMIT = 1u << 22;
// Where a trap came from: a host-mapped word (set by the host through ccsp::tc::set_trap_info_ptr) that the trapping thread
// fills in first — it survives the death of the CUDA context, so the host can still read it (ccsp_debug_trap_info()).
__device__ unsigned long long *g_trap_info = nullptr;
__device__ __noinline__ void trap_with(unsigned code) {
  if (g_trap_info) {
    g_trap_info[0] = ((unsigned long long)code << 40) | ((unsigned long long)(blockIdx.x & 0xFFFFu) << 24) | (threadIdx.x & 0xFFFFFFu);
    __threadfence_system();
  }
  __trap();
}
#ifdef CCSP_DEBUG_SPIN_TRAP
#define CCSP_SPIN_GUARD(spins) do { if (++(spins) > ::ccsp::tc::SPIN_LIMIT) ::ccsp::tc::trap_with(__LINE__); } while (0)
#else
#define CCSP_SPIN_GUARD(spins) do { (void)(spins); } while (0)
#endif

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
    CCSP_SPIN_GUARD(spins);
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// multicast variant: the bytes and the complete_tx land at the same CTA-relative offsets in every CTA of `mask`
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void *src, uint32_t bytes, uint64_t *bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// src_bytes = 0: the 16 destination bytes are zero-filled and NO global read is issued (padded edge rows: thousands of them
// would otherwise hammer the one 1 KB zero row of pe_split from every CTA — an L2 same-line hot spot)
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void *src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

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

// arrives on the barrier at the same CTA-relative offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

__device__ __forceinline__ float4 ldg_nc_f4(const float4 *p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
// SiLU with the hardware exp2/rcp approximations (|error| <~ 3e-7 |silu(x)| + 1e-7; measured parity impact
// nil); the FP32 validation path keeps the libm-accurate version (common.cuh silu_f).
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
// SWIZZLE_128B K-major tile (rows of 128 B, 8-row groups of 1024 B): layout type 2, SBO = 1024 B.  The start address may be
// advanced by multiples of 32 B inside the 128-byte row (one K = 16 BF16 step each): the XOR acts on absolute address bits.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// byte offset of 16-byte piece `q` (0..7) of row `r` inside a SWIZZLE_128B operand tile (Swizzle<3,4,3>)
__host__ __device__ __forceinline__ uint32_t sw128_off(uint32_t r, uint32_t q) { return r * 128 + ((q ^ (r & 7)) << 4); }
// byte offset of 16-byte piece `q` (0..3) of row `r` inside a SWIZZLE_64B operand tile (Swizzle<2,4,3>)
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
// arithmetic mode
// ---------------------------------------------------------------------------------------------------
template <int KIND_, int NSPLIT_>
struct Mode {
  static constexpr int KIND = KIND_, NSPLIT = NSPLIT_;
  static constexpr int ELT = KIND == KIND_TF32 ? 4 : 2;      // operand element bytes
  static constexpr int KC = ROWB / ELT;                      // K elements per stage (16 / 32)
  static constexpr int KSTEPS = 2;                           // tcgen05.mma K (8 / 16 elements = 32 B) per stage
  static constexpr int NS = NSPLIT == 3 ? 2 : 1;             // operand parts per stage (hi[, lo])
  static constexpr int A_STAGE = NS * PART;
  // split pose embedding row written by the node kernel.
  //   TF32: [256 hi][256 lo] operand words.
  //   BF16: per 64-byte k-chunk (32 elements) [hi 64 B | lo 64 B], 8 chunks: the hi and lo slices one ring stage needs
  //         from a gathered row are 128 contiguous bytes = one cache line and one shared-memory wavefront per row,
  //         instead of two half lines (the 16-byte cp.async gather costs a wavefront per distinct global segment).
  static constexpr int PE_ROW_BYTES = 2 * CCSP_H * ELT;
  // byte offset, inside a row, of 16-byte piece q (0..3) of operand part `part` (0 hi, 1 lo) of k-chunk kc (of the 256 dims)
  __host__ __device__ static constexpr int pe_off(int kc, int part, int q) {
    return KIND == KIND_TF32 ? part * (CCSP_H * ELT) + kc * ROWB + q * 16 : kc * (2 * ROWB) + part * ROWB + q * 16;
  }
  // byte offset of element k (0..255) of part `part`
  __host__ __device__ static constexpr int pe_elem_off(int k, int part) {
    return pe_off(k / KC, part, (k % KC) * ELT / 16) + ((k % KC) * ELT) % 16;
  }
  static constexpr int NKC1 = CCSP_H2 / KC;                  // first layer: K = 512
  static constexpr int NKC2 = CCSP_H / KC;                   // decoder:     K = 256
  // bytes of H (operand format) per 128-row x 256-K decoder tile
  static constexpr size_t H_TILE_BYTES = (size_t)NKC2 * A_STAGE;
  __host__ __device__ static constexpr uint32_t idesc(int n) {
    return (1u << 4) | ((KIND == KIND_TF32 ? 2u : 1u) << 7) | ((KIND == KIND_TF32 ? 2u : 1u) << 10) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(SUB_M >> 4) << 24);
  }
};

// one k-chunk of MMAs: D[128 x N] (+)= A_stage . B_stage^T
template <class M, int NTILE>
__device__ __forceinline__ void issue_chunk(uint32_t d_tmem, uint32_t a_hi, uint32_t b_hi, bool first) {
  const uint32_t a_lo = a_hi + PART, b_lo = b_hi + NTILE * ROWB;
#pragma unroll
  for (int ks = 0; ks < M::KSTEPS; ++ks) {
    const uint32_t ko = ks * 32;
    if (M::NSPLIT == 3) {
      umma<M::KIND>(d_tmem, smem_desc_sw64(a_lo + ko), smem_desc_sw64(b_hi + ko), M::idesc(NTILE), !(first && ks == 0));
      umma<M::KIND>(d_tmem, smem_desc_sw64(a_hi + ko), smem_desc_sw64(b_lo + ko), M::idesc(NTILE), 1);
      umma<M::KIND>(d_tmem, smem_desc_sw64(a_hi + ko), smem_desc_sw64(b_hi + ko), M::idesc(NTILE), 1);
    } else {
      umma<M::KIND>(d_tmem, smem_desc_sw64(a_hi + ko), smem_desc_sw64(b_hi + ko), M::idesc(NTILE), !(first && ks == 0));
    }
  }
}

// store 32 consecutive FP32 values of one row as operand pieces (hi[, lo]) into a SWIZZLE_64B tile image
//   TF32: two k-chunks of 16 values;  BF16: one k-chunk of 32 values.  `base` points at the first k-chunk,
//   `chunk_stride` is the byte distance between consecutive k-chunks (A_STAGE).
template <class M>
__device__ __forceinline__ void store_split32(uint8_t *base, size_t chunk_stride, int r, const float *h) {
  if (M::KIND == KIND_TF32) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float *f = h + c * 16 + q * 4;
        uint4 hi;
        hi.x = tf32_rna(f[0]); hi.y = tf32_rna(f[1]); hi.z = tf32_rna(f[2]); hi.w = tf32_rna(f[3]);
        uint8_t *dst = base + c * chunk_stride + sw64_off(r, q);
        *reinterpret_cast<uint4 *>(dst) = hi;
        if (M::NS == 2) {
          uint4 lo;
          lo.x = tf32_rna(f[0] - __uint_as_float(hi.x)); lo.y = tf32_rna(f[1] - __uint_as_float(hi.y));
          lo.z = tf32_rna(f[2] - __uint_as_float(hi.z)); lo.w = tf32_rna(f[3] - __uint_as_float(hi.w));
          *reinterpret_cast<uint4 *>(dst + PART) = lo;
        }
      }
    }
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float *f = h + q * 8;
      float g[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = __bfloat162float(__float2bfloat16_rn(f[i]));
      uint4 hi;
      hi.x = pack_bf16(g[0], g[1]); hi.y = pack_bf16(g[2], g[3]); hi.z = pack_bf16(g[4], g[5]); hi.w = pack_bf16(g[6], g[7]);
      uint8_t *dst = base + sw64_off(r, q);
      *reinterpret_cast<uint4 *>(dst) = hi;
      if (M::NS == 2) {
        uint4 lo;
        lo.x = pack_bf16(f[0] - g[0], f[1] - g[1]); lo.y = pack_bf16(f[2] - g[2], f[3] - g[3]);
        lo.z = pack_bf16(f[4] - g[4], f[5] - g[5]); lo.w = pack_bf16(f[6] - g[6], f[7] - g[7]);
        *reinterpret_cast<uint4 *>(dst + PART) = lo;
      }
    }
  }
}

// Same split, but for a GLOBAL destination: each row's 64 B of an operand part are written as two 32-byte
// stores (st.global.v8.b32 -> STG.256), i.e. full 32-byte sectors.  With 16-byte stores every sector of the
// SWIZZLE_64B image is written in two halves by two different instructions, which made the H stores the most
// expensive part of the first-layer epilogue (ablation in profiles/README.md).
__device__ __forceinline__ void stg256(uint8_t *dst, const uint4 &a, const uint4 &b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(dst), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void swap_u4(uint4 &a, uint4 &b, bool c) {
  uint4 t = a;
  a.x = c ? b.x : a.x; a.y = c ? b.y : a.y; a.z = c ? b.z : a.z; a.w = c ? b.w : a.w;
  b.x = c ? t.x : b.x; b.y = c ? t.y : b.y; b.z = c ? t.z : b.z; b.w = c ? t.w : b.w;
}
// pieces[q] (logical 16-byte pieces of one row) -> physical order of the swizzled row, two 32-byte stores
__device__ __forceinline__ void store_row64_swz(uint8_t *row_base, uint4 *pc, int r) {
  const int sw = (r >> 1) & 3;
  swap_u4(pc[0], pc[1], sw & 1); swap_u4(pc[2], pc[3], sw & 1);
  swap_u4(pc[0], pc[2], sw & 2); swap_u4(pc[1], pc[3], sw & 2);
  stg256(row_base, pc[0], pc[1]);
  stg256(row_base + 32, pc[2], pc[3]);
}
template <class M>
__device__ __forceinline__ void store_split32_global(uint8_t *base, size_t chunk_stride, int r, const float *h) {
  if (M::KIND == KIND_TF32) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint4 hi[4], lo[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float *f = h + c * 16 + q * 4;
        hi[q].x = tf32_rna(f[0]); hi[q].y = tf32_rna(f[1]); hi[q].z = tf32_rna(f[2]); hi[q].w = tf32_rna(f[3]);
        lo[q].x = tf32_rna(f[0] - __uint_as_float(hi[q].x)); lo[q].y = tf32_rna(f[1] - __uint_as_float(hi[q].y));
        lo[q].z = tf32_rna(f[2] - __uint_as_float(hi[q].z)); lo[q].w = tf32_rna(f[3] - __uint_as_float(hi[q].w));
      }
      uint8_t *rb = base + c * chunk_stride + (size_t)r * ROWB;
      store_row64_swz(rb, hi, r);
      if (M::NS == 2) store_row64_swz(rb + PART, lo, r);
    }
  } else {
    uint4 hi[4], lo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float *f = h + q * 8;
      float g[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = __bfloat162float(__float2bfloat16_rn(f[i]));
      hi[q].x = pack_bf16(g[0], g[1]); hi[q].y = pack_bf16(g[2], g[3]); hi[q].z = pack_bf16(g[4], g[5]); hi[q].w = pack_bf16(g[6], g[7]);
      lo[q].x = pack_bf16(f[0] - g[0], f[1] - g[1]); lo[q].y = pack_bf16(f[2] - g[2], f[3] - g[3]);
      lo[q].z = pack_bf16(f[4] - g[4], f[5] - g[5]); lo[q].w = pack_bf16(f[6] - g[6], f[7] - g[7]);
    }
    uint8_t *rb = base + (size_t)r * ROWB;
    store_row64_swz(rb, hi, r);
    if (M::NS == 2) store_row64_swz(rb + PART, lo, r);
  }
}

// ===================================================================================================
// first layer
// ===================================================================================================
struct L1Args {
  const uint8_t *pe_split;   // [(n+1)][PE_ROW_BYTES] split pose embeddings (row n = zeros)
  const int *src_i, *src_j;  // [Epad] edge endpoints (padded rows -> n)
  const uint8_t *b_blob;     // [C][2][NKC1][B_STAGE] packed pose-column weights
  const int *tile_type;      // [Epad/128] constraint type per tile
  int num_m_tiles;
  const float *S;            // static term, blocked layout (common.cuh blk_off)
  const float *tb;           // [C][512] time term at the current timestep
  uint8_t *H;                // operand-format activations: [tile][slot][NKC2][NS][PART]
  int dbg;                   // harness ablations: 1 no A gather, 2 no B copy, 4 no epilogue IO (16 loads only, 32 stores only), 8 no MMA, 64 drain only
};

// CL = thread-block-cluster size.  The CL CTAs of a cluster work on CL consecutive 128-edge tiles of the
// same constraint type and the same slot, so they need the SAME weight stage at the same time: each CTA fetches
// 1/CL of it and multicasts that piece into every CTA's ring (cp.async.bulk ... .multicast::cluster), cutting
// the L2->SM weight traffic — the binding resource of this kernel — by CL.  A stage is recycled only when every
// CTA of the cluster has consumed it: the MMA warps commit to the `empty` barrier of ALL CTAs (count = CL).
template <class M, int CL_ = 1>
struct L1Cfg {
  static constexpr int CL = CL_;
  static constexpr int NTILE = 256;
  static constexpr int B_STAGE = M::NS * NTILE * ROWB;
  static constexpr int STAGE = M::A_STAGE + B_STAGE;
  static constexpr int NSTAGE = 4;
  static constexpr int LAG = 2;                              // cp.async chunks in flight per producer thread
  static constexpr int NBUF = 2;                             // TMEM accumulator buffers (2 x 256 columns)
  static constexpr int WARP_PROD0 = NUM_EPI_WARPS, NUM_PROD_WARPS = 4;
  static constexpr int WARP_LOAD = WARP_PROD0 + NUM_PROD_WARPS, WARP_MMA = WARP_LOAD + 1;
  static constexpr int THREADS = (WARP_MMA + 1) * 32;        // 448
  static constexpr int SMEM_EXTRA = 2048;
  static constexpr int SMEM_BYTES = NSTAGE * STAGE + SMEM_EXTRA + 1024;
};

template <class M, int CL>
__global__ void __launch_bounds__(L1Cfg<M, CL>::THREADS, 1) k_edge_l1_tc(const L1Args A) {
  using C = L1Cfg<M, CL>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *extra = smem + C::NSTAGE * C::STAGE;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(extra);            // [NSTAGE]
  uint64_t *empty_bar = full_bar + 8;                                  // [NSTAGE]
  uint64_t *tfull_bar = empty_bar + 8;                                 // [NBUF]
  uint64_t *tempty_bar = tfull_bar + 4;                                // [NBUF]
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tempty_bar + 4);
  float *tb_s = reinterpret_cast<float *>(extra + 512);                // [256] time-term slice of the current tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work units: (group of CL consecutive 128-edge tiles, slot); this CTA handles tile  group * CL + rank
  const int num_units = (A.num_m_tiles / CL) * 2;
  const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int unit0 = blockIdx.x / CL, unit_step = gridDim.x / CL;
  const uint16_t mc_mask = (uint16_t)((1u << CL) - 1);
  const uint32_t smem_base = smem_u32(smem);
#define L1_TILE_OF(u) ((((u) >> 1) * CL + rank) * 2 + ((u) & 1))

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::NSTAGE; ++s) { mbar_init(&full_bar[s], C::NUM_PROD_WARPS * 32 + 1); mbar_init(&empty_bar[s], CL); }
    for (int b = 0; b < C::NBUF; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], EPI_THREADS); }
    fence_barrier_init();
  }
  if (warp == C::WARP_MMA) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync();          // every CTA's barriers are initialised before any peer signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp >= C::WARP_PROD0 && warp < C::WARP_LOAD) {
    // ============ A gather: 128 threads, thread = (piece q, rows r0 + 32 p); cp.async, LAG chunks in flight
    const int t = threadIdx.x - C::WARP_PROD0 * 32;
    const int q = t & 3, r0 = t >> 2;
    uint32_t g = 0;                      // global chunk counter (issue side)
    uint32_t sig = 0;                    // chunks already signalled full
    for (int u = unit0; u < num_units; u += unit_step) {
      const int tile = L1_TILE_OF(u);
      const int m0 = (tile >> 1) * SUB_M;
      size_t roff[4];
#pragma unroll 1
      for (int kc = 0; kc < M::NKC1; ++kc, ++g) {
        if (kc == 0 || kc == M::NKC1 / 2) {
          const int *idx = kc == 0 ? A.src_i : A.src_j;
#pragma unroll
          for (int p = 0; p < 4; ++p) roff[p] = (size_t)__ldg(&idx[m0 + r0 + 32 * p]) * M::PE_ROW_BYTES;
        }
        const uint32_t s = g % C::NSTAGE;
        mbar_wait(&empty_bar[s], ((g / C::NSTAGE) & 1) ^ 1);
        if (!(A.dbg & 1)) {
          const int kcl = kc % (M::NKC1 / 2);
          const uint32_t st = smem_base + s * C::STAGE;
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const uint8_t *src = A.pe_split + roff[p];
            const uint32_t dst = st + sw64_off(r0 + 32 * p, q);
            cp_async16(dst, src + M::pe_off(kcl, 0, q));
            if (M::NS == 2) cp_async16(dst + PART, src + M::pe_off(kcl, 1, q));
          }
        }
        cp_async_commit();
        if (g - sig >= (uint32_t)C::LAG) {         // chunk `sig` has landed: publish it to the async proxy
          cp_async_wait<C::LAG>();
          fence_proxy_async();
          mbar_arrive(&full_bar[sig % C::NSTAGE]);
          ++sig;
        }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    for (; sig < g; ++sig) mbar_arrive(&full_bar[sig % C::NSTAGE]);
  } else if (warp == C::WARP_LOAD) {
    // ============ B loader ======================================================================
    if (lane == 0) {
      uint32_t g = 0;
      constexpr uint32_t PIECE = C::B_STAGE / CL;
      for (int u = unit0; u < num_units; u += unit_step) {
        const int tile = L1_TILE_OF(u);
        const int grp = __ldg(&A.tile_type[tile >> 1]);
        const uint8_t *blob = A.b_blob + ((size_t)(grp * 2 + (tile & 1)) * M::NKC1) * C::B_STAGE;
        for (int kc = 0; kc < M::NKC1; ++kc, ++g) {
          const uint32_t s = g % C::NSTAGE;
          mbar_wait(&empty_bar[s], ((g / C::NSTAGE) & 1) ^ 1);       // all CL CTAs have consumed this stage
          if (A.dbg & 2) { mbar_arrive(&full_bar[s]); continue; }
          mbar_arrive_expect_tx(&full_bar[s], C::B_STAGE);           // own piece + the peers' multicast pieces
          const uint32_t dst = smem_base + s * C::STAGE + M::A_STAGE + rank * PIECE;
          const uint8_t *src = blob + (size_t)kc * C::B_STAGE + rank * PIECE;
          if (CL > 1) bulk_g2s_mc(dst, src, PIECE, &full_bar[s], mc_mask);
          else bulk_g2s(dst, src, PIECE, &full_bar[s]);
        }
      }
    }
  } else if (warp == C::WARP_MMA) {
    // ============ MMA issuer ====================================================================
    if (lane == 0) {
      uint32_t g = 0, tcount = 0;
      for (int u = unit0; u < num_units; u += unit_step, ++tcount) {
        const uint32_t buf = tcount % C::NBUF;
        mbar_wait(&tempty_bar[buf], ((tcount / C::NBUF) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * C::NTILE;
        for (int kc = 0; kc < M::NKC1; ++kc, ++g) {
          const uint32_t s = g % C::NSTAGE;
          mbar_wait(&full_bar[s], (g / C::NSTAGE) & 1);
          tc_fence_after();
          const uint32_t a_hi = smem_base + s * C::STAGE;
          if (!(A.dbg & 8)) issue_chunk<M, C::NTILE>(d_tmem, a_hi, a_hi + M::A_STAGE, kc == 0);
          if (CL > 1) umma_commit_mc(&empty_bar[s], mc_mask);   // stage consumed: tell every CTA of the cluster
          else umma_commit(&empty_bar[s]);
        }
        umma_commit(&tfull_bar[buf]);        // accumulator complete -> epilogue
      }
    }
  } else if (warp < NUM_EPI_WARPS) {
    // ============ epilogue: warp w <-> TMEM lanes 32 (w & 3).., columns 128 (w >> 2).. +127 ========
    uint32_t tcount = 0;
    const int quarter = warp & 3, chalf = warp >> 2;
    const int r = quarter * 32 + lane;                                  // row within the tile
    for (int u = unit0; u < num_units; u += unit_step, ++tcount) {
      const int tile = L1_TILE_OF(u);
      const uint32_t buf = tcount % C::NBUF;
      const int mt = tile >> 1, slot = tile & 1;
      const int grp = __ldg(&A.tile_type[mt]);
      const size_t row = (size_t)mt * SUB_M + r;
      const int col0 = slot * 256 + chalf * 128;                        // first of this thread's 128 columns
      // S in the blocked layout: 32-row x 32-col blocks of 8 pieces x 32 lanes x 16 B  (coalesced 512 B / instr)
      const float4 *Sblk = reinterpret_cast<const float4 *>(A.S) + (((row >> 5) * 16 + (col0 >> 5)) * 8) * 32 + lane;
      float4 sn[8];                                                    // prefetched one 32-column chunk ahead
#pragma unroll
      for (int j = 0; j < 8; ++j) sn[j] = (A.dbg & (4 | 16)) ? make_float4(0.f, 0.f, 0.f, 0.f) : ldg_nc_f4(Sblk + j * 32);
      const float tb_mine = __ldg(&A.tb[(size_t)grp * CCSP_H2 + slot * 256 + threadIdx.x]);
      asm volatile("bar.sync 1, 256;" ::: "memory");                  // previous tile's readers of tb_s are done
      tb_s[threadIdx.x] = tb_mine;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(&tfull_bar[buf], (tcount / C::NBUF) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * C::NTILE + chalf * 128 + ((uint32_t)(quarter * 32) << 16);
      uint8_t *Hrow = A.H + (size_t)tile * M::H_TILE_BYTES + (size_t)(chalf * 128 / M::KC) * M::A_STAGE;
#pragma unroll 1
      for (int cb = 0; cb < 128; cb += 32) {
        float4 sc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) sc[j] = sn[j];
        if (cb + 32 < 128 && !(A.dbg & (4 | 16))) {
#pragma unroll
          for (int j = 0; j < 8; ++j) sn[j] = ldg_nc_f4(Sblk + ((cb + 32) >> 5) * 256 + j * 32);
        }
        float v[32];
        tmem_ld32(taddr + cb, v);
        if (A.dbg & 64) continue;            // ablation: TMEM drain only
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t4 = *reinterpret_cast<const float4 *>(&tb_s[chalf * 128 + cb + 4 * j]);
          v[4 * j] = silu_fast(v[4 * j] + sc[j].x + t4.x);
          v[4 * j + 1] = silu_fast(v[4 * j + 1] + sc[j].y + t4.y);
          v[4 * j + 2] = silu_fast(v[4 * j + 2] + sc[j].z + t4.z);
          v[4 * j + 3] = silu_fast(v[4 * j + 3] + sc[j].w + t4.w);
        }
        if (!(A.dbg & (4 | 32))) store_split32_global<M>(Hrow + (size_t)(cb / M::KC) * M::A_STAGE, M::A_STAGE, r, v);
        else if (v[0] == 12345.678f && v[31] == 9.f) Hrow[0] = 1;     // keep the math alive in the ablations
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync();          // no CTA exits while a peer may still multicast into it
  if (warp == C::WARP_MMA) tmem_dealloc(tmem_base, 512);
#undef L1_TILE_OF
}

// ===================================================================================================
// decoder
// ===================================================================================================
struct DecArgs {
  const uint8_t *H;          // operand-format activations written by k_edge_l1_tc
  const uint8_t *b_blob;     // [NKC2][B_STAGE] packed pose_decoder.0 weights
  int num_tiles;             // (128-edge tile, slot) pairs = 2 * Epad / 128
  const float *bd1, *Wd2, *bd2;
  int P;
  float *o;                  // [Epad][2][P]
  int dbg;
};

template <class M>
struct DecCfg {
  static constexpr int NTILE = 128;
  static constexpr int B_STAGE = M::NS * NTILE * ROWB;
  static constexpr int STAGE = M::A_STAGE + B_STAGE;
  static constexpr int NSTAGE = 6;
  static constexpr int NBUF = 4;                             // 4 x 128 TMEM columns
  static constexpr int WARP_LOAD = NUM_EPI_WARPS, WARP_MMA = WARP_LOAD + 1;
  static constexpr int THREADS = (WARP_MMA + 1) * 32;        // 320
  static constexpr int SMEM_EXTRA = 512 + (CCSP_HH + CCSP_MAXP * CCSP_HH + CCSP_MAXP) * 4 + SUB_M * CCSP_MAXP * 4;
  static constexpr int SMEM_BYTES = NSTAGE * STAGE + SMEM_EXTRA + 1024;
};

template <class M>
__global__ void __launch_bounds__(DecCfg<M>::THREADS, 1) k_edge_dec_tc(const DecArgs A) {
  using C = DecCfg<M>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *extra = smem + C::NSTAGE * C::STAGE;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(extra);            // [NSTAGE]
  uint64_t *empty_bar = full_bar + 8;
  uint64_t *tfull_bar = empty_bar + 8;                                 // [NBUF]
  uint64_t *tempty_bar = tfull_bar + 4;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tempty_bar + 4);
  float *bd1 = reinterpret_cast<float *>(extra + 512);                 // [128]
  float *w2t = bd1 + CCSP_HH;                                          // [128][8]  (transposed, zero-padded)
  float *bd2 = w2t + CCSP_MAXP * CCSP_HH;                              // [8]
  float *red = bd2 + CCSP_MAXP;                                        // [128][8] partial sums of the upper column half

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = smem_u32(smem);

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < C::NBUF; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], EPI_THREADS); }
    fence_barrier_init();
  }
  if (warp == C::WARP_MMA) tmem_alloc(tmem_ptr, 512);
  if (warp < NUM_EPI_WARPS) {
    for (int i = threadIdx.x; i < CCSP_HH; i += EPI_THREADS) bd1[i] = A.bd1[i];
    for (int i = threadIdx.x; i < CCSP_MAXP * CCSP_HH; i += EPI_THREADS) {
      const int j = i / CCSP_MAXP, pp = i % CCSP_MAXP;
      w2t[i] = pp < A.P ? A.Wd2[pp * CCSP_HH + j] : 0.f;
    }
    if (threadIdx.x < CCSP_MAXP) bd2[threadIdx.x] = threadIdx.x < A.P ? A.bd2[threadIdx.x] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == C::WARP_LOAD) {
    if (lane == 0) {
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < A.num_tiles; tile += gridDim.x) {
        const uint8_t *a = A.H + (size_t)tile * M::H_TILE_BYTES;
        for (int kc = 0; kc < M::NKC2; ++kc, ++g) {
          const uint32_t s = g % C::NSTAGE;
          mbar_wait(&empty_bar[s], ((g / C::NSTAGE) & 1) ^ 1);
          if (A.dbg & 3) { mbar_arrive(&full_bar[s]); continue; }
          mbar_arrive_expect_tx(&full_bar[s], C::STAGE);
          bulk_g2s(smem_base + s * C::STAGE, a + (size_t)kc * M::A_STAGE, M::A_STAGE, &full_bar[s]);
          bulk_g2s(smem_base + s * C::STAGE + M::A_STAGE, A.b_blob + (size_t)kc * C::B_STAGE, C::B_STAGE, &full_bar[s]);
        }
      }
    }
  } else if (warp == C::WARP_MMA) {
    if (lane == 0) {
      uint32_t g = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < A.num_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t buf = tcount % C::NBUF;
        mbar_wait(&tempty_bar[buf], ((tcount / C::NBUF) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * C::NTILE;
        for (int kc = 0; kc < M::NKC2; ++kc, ++g) {
          const uint32_t s = g % C::NSTAGE;
          mbar_wait(&full_bar[s], (g / C::NSTAGE) & 1);
          tc_fence_after();
          const uint32_t a_hi = smem_base + s * C::STAGE;
          if (!(A.dbg & 8)) issue_chunk<M, C::NTILE>(d_tmem, a_hi, a_hi + M::A_STAGE, kc == 0);
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tfull_bar[buf]);
      }
    }
  } else if (warp < NUM_EPI_WARPS) {
    // warp w <-> TMEM lanes 32 (w & 3).., columns 64 (w >> 2).. +63; the two column halves of a row are
    // combined through shared memory (fixed order: lower half + upper half) by the lower-half warp.
    uint32_t tcount = 0;
    const int quarter = warp & 3, chalf = warp >> 2;
    const int r = quarter * 32 + lane;
    for (int tile = blockIdx.x; tile < A.num_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t buf = tcount % C::NBUF;
      mbar_wait(&tfull_bar[buf], (tcount / C::NBUF) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * C::NTILE + chalf * 64 + ((uint32_t)(quarter * 32) << 16);
      float acc[CCSP_MAXP];
#pragma unroll
      for (int p = 0; p < CCSP_MAXP; ++p) acc[p] = 0.f;
#pragma unroll 1
      for (int cb = 0; cb < 64; cb += 32) {
        float v[32];
        tmem_ld32(taddr + cb, v);
        const int c0 = chalf * 64 + cb;
        if (A.P <= 4) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = silu_fast(v[j] + bd1[c0 + j]);
            const float4 w = *reinterpret_cast<const float4 *>(&w2t[(c0 + j) * CCSP_MAXP]);
            acc[0] = fmaf(d, w.x, acc[0]); acc[1] = fmaf(d, w.y, acc[1]); acc[2] = fmaf(d, w.z, acc[2]); acc[3] = fmaf(d, w.w, acc[3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = silu_fast(v[j] + bd1[c0 + j]);
            const float4 w0 = *reinterpret_cast<const float4 *>(&w2t[(c0 + j) * CCSP_MAXP]);
            const float4 w1 = *reinterpret_cast<const float4 *>(&w2t[(c0 + j) * CCSP_MAXP + 4]);
            acc[0] = fmaf(d, w0.x, acc[0]); acc[1] = fmaf(d, w0.y, acc[1]); acc[2] = fmaf(d, w0.z, acc[2]); acc[3] = fmaf(d, w0.w, acc[3]);
            acc[4] = fmaf(d, w1.x, acc[4]); acc[5] = fmaf(d, w1.y, acc[5]); acc[6] = fmaf(d, w1.z, acc[6]); acc[7] = fmaf(d, w1.w, acc[7]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[buf]);                                   // accumulator drained: MMA may reuse it
      if (chalf == 1) {
        *reinterpret_cast<float4 *>(&red[r * CCSP_MAXP]) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        if (A.P > 4) *reinterpret_cast<float4 *>(&red[r * CCSP_MAXP + 4]) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (chalf == 0 && !(A.dbg & 4)) {
        const int mt = tile >> 1, slot = tile & 1;
        float *orow = A.o + ((size_t)(mt * SUB_M + r) * 2 + slot) * A.P;
        float res[CCSP_MAXP];
#pragma unroll
        for (int p = 0; p < CCSP_MAXP; ++p) res[p] = (acc[p] + red[r * CCSP_MAXP + p]) + bd2[p];
        if (A.P == 4) {
          *reinterpret_cast<float4 *>(orow) = make_float4(res[0], res[1], res[2], res[3]);
        } else {
#pragma unroll
          for (int p = 0; p < CCSP_MAXP; ++p)
            if (p < A.P) orow[p] = res[p];
        }
      }
      asm volatile("bar.sync 3, 256;" ::: "memory");                  // `red` may be overwritten by the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == C::WARP_MMA) tmem_dealloc(tmem_base, 512);
}

// ===================================================================================================
// arguments of the fused edge kernel (BF16 operand modes, kernels_fused2.cuh):
// first layer -> SiLU -> split -> decoder -> o, H never leaves the SM
// ===================================================================================================
struct FusedArgs {
  const uint8_t *pe_split;   // [(n+1)][PE_ROW_BYTES]
  const int *src_i, *src_j;  // [Epad]
  const uint8_t *b_blob;     // [C][2][NKC1][B1_STAGE] first-layer pose-column weights
  const uint8_t *w_blob;     // [NKC2][B2_STAGE] pose_decoder.0 weights
  const int *tile_type;
  int num_m_tiles;
  const float *S;            // blocked layout
  const float *tb;           // [C][512] at the current timestep
  const float *bd1, *Wd2, *bd2;
  int P;
  float *o;                  // [Epad][2][P]
  int dbg;
  long long *trace;          // harness only: clock64 timeline of CTA 0 ([role][unit < 8][16]), else nullptr
  int pad_row_plus1;         // 1 + index of the zero row that padded edges point at (0: unknown, gather it like any row)
  // persistent mode (k_edge_fused2_tc<.., PERSIST = true>): the kernel stays resident for ALL evaluations of a sample() next to
  // the persistent node kernel and hand-shakes with it through two device counters instead of kernel boundaries
  int num_evals;             // denoiser evaluations of the sample
  const int *eval_t;         // [num_evals] timestep of each evaluation: tb = tb_base + eval_t[ev] * tb_stride
  const float *tb_base;
  int tb_stride;
  unsigned *node_done;       // += 1 per node CTA and node iteration; evaluation ev may gather once it reaches (ev + 1) * node_ctas
  unsigned *edge_done;       // += 1 per edge CTA and evaluation (all of its o rows written)
  unsigned node_ctas;
  // pipelined chains (persistent mode only): the plan's tiles are grouped into num_chains independent scene groups ("chains");
  // chain c owns tiles chain_tile0[c] .. chain_tile0[c + 1] and the flag words node_done + 32 c / edge_done + 32 c.  The
  // kernel runs (evaluation 0, chain 0), (0, 1), .., (1, 0), ..: while the node kernel updates the nodes of one chain, the
  // tensor pipe works on another chain, and the operand pipeline never drains.  num_chains <= 1: one chain = all tiles.
  int num_chains;
  int chain_tile0[CCSP_MAX_CHAINS + 1];
  unsigned chain_nblk[CCSP_MAX_CHAINS];   // node blocks of chain c: with chains, node_done counts BLOCKS (a chain evaluation may
                             // gather at (ev + 1) * chain_nblk[c]) and edge_done counts units x 2 CTAs, so a CTA without work in a chain
                             // evaluation neither waits nor signals and the chains only meet on the SMs they share
  int drain_each_eval;       // 1: epilogue-2 is drained at the end of every chain evaluation (0: carried into the next one)
  unsigned *arrive;          // host-mapped counter, += 1 per CTA at entry (the host launches the node kernel once all are resident)
};

// ---- device-scope flags of the persistent pair of kernels ------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned *p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// persistent-mode timeline (developer aid, CCSP_PERSIST_TRACE=1): trace[event][iteration < 32] in ns of the global timer
#define PTRACE(buf, event, iter) do { if ((buf) && (iter) < 32) (buf)[(event) * 32 + (iter)] = ::ccsp::tc::global_ns(); } while (0)
// Bounded on purpose, in every build: the two persistent kernels must be co-resident; if they are not (the launcher's
// occupancy rule was violated) a trap is the only alternative to a device-wide hang.  2^27 polls x >= 100 ns > 10 s.
__device__ __forceinline__ void wait_flag_ge(const unsigned *p, unsigned target) {
  unsigned spins = 0;
  while ((int)(ld_acquire_gpu_u32(p) - target) < 0) {
    if (++spins > (1u << 27)) trap_with(900000u + (target & 0xFFFFu));
    __nanosleep(40);
  }
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

template <class M, int NTILE>
void pack_b_blob(const float *W, int ldw, int k_begin, int K_total, int n_rows_total, uint8_t *out) {
  const int n_tiles = n_rows_total / NTILE, NKC = K_total / M::KC;
  const int EPC = 16 / M::ELT;   // elements per 16-byte piece
  const int B_PART = NTILE * ROWB, B_STAGE = M::NS * B_PART;
  for (int nt = 0; nt < n_tiles; ++nt)
    for (int kc = 0; kc < NKC; ++kc) {
      uint8_t *st = out + ((size_t)nt * NKC + kc) * B_STAGE;
      for (int r = 0; r < NTILE; ++r)
        for (int kk = 0; kk < M::KC; ++kk) {
          const float x = W[(size_t)(nt * NTILE + r) * ldw + k_begin + kc * M::KC + kk];
          const uint32_t off = sw64_off(r, kk / EPC) + (kk % EPC) * M::ELT;
          if (M::KIND == KIND_TF32) {
            uint32_t hi = host_tf32_rna(x);
            float hf;
            memcpy(&hf, &hi, 4);
            memcpy(st + off, &hi, 4);
            if (M::NS == 2) { uint32_t lo = host_tf32_rna(x - hf); memcpy(st + B_PART + off, &lo, 4); }
          } else {
            uint16_t hi = host_bf16_rn(x);
            memcpy(st + off, &hi, 2);
            if (M::NS == 2) { uint16_t lo = host_bf16_rn(x - host_bf16_to_f(hi)); memcpy(st + B_PART + off, &lo, 2); }
          }
        }
    }
}

template <class M, int CL = 1>
cudaError_t launch_l1_tc(const L1Args &a, int num_sms, cudaStream_t st) {
  using C = L1Cfg<M, CL>;
  static int max_clusters_dev[64] = {};      // per device: function attributes and cluster occupancy
  int dev_ = 0;
  cudaGetDevice(&dev_);
  int &max_clusters = max_clusters_dev[dev_ & 63];
  if (max_clusters == 0) {
    cudaError_t e = cudaFuncSetAttribute(k_edge_l1_tc<M, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    max_clusters = num_sms / CL;
    if (CL > 1) {           // how many clusters can be co-resident (GPC granularity can strand SMs)
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(num_sms / CL * CL); q.blockDim = dim3(C::THREADS); q.dynamicSmemBytes = C::SMEM_BYTES;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      q.attrs = at; q.numAttrs = 1;
      int n = 0;
      e = cudaOccupancyMaxActiveClusters(&n, k_edge_l1_tc<M, CL>, &q);
      if (e != cudaSuccess) return e;
      if (n > 0 && n < max_clusters) max_clusters = n;
    }
  }
  if (a.num_m_tiles % CL != 0) return cudaErrorInvalidValue;
  const int units = (a.num_m_tiles / CL) * 2;
  if (units == 0) return cudaSuccess;
  const int nclusters = units < max_clusters ? units : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nclusters * CL); cfg.blockDim = dim3(C::THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = CL; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k_edge_l1_tc<M, CL>, a);
}

template <class M>
cudaError_t launch_dec_tc(const DecArgs &a, int num_sms, cudaStream_t st) {
  using C = DecCfg<M>;
  static bool configured_dev[64] = {};      // per device: function attributes belong to the device's context
  int dev_ = 0;
  cudaGetDevice(&dev_);
  bool &configured = configured_dev[dev_ & 63];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_edge_dec_tc<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (a.num_tiles == 0) return cudaSuccess;
  k_edge_dec_tc<M><<<a.num_tiles < num_sms ? a.num_tiles : num_sms, C::THREADS, C::SMEM_BYTES, st>>>(a);
  return cudaGetLastError();
}

}  // namespace tc
}  // namespace ccsp
