// kernels_fused2.cuh — the per-edge MLP as ONE kernel on a CTA pair (tcgen05 cta_group::2), BF16 operand modes.
//
//   o[e, slot, :] = pose_decoder( SiLU( [pe_i ; pe_j] W_c,pose^T + S[e] + tb[t,c] )[slot half] )
//   (denoise_fn.py:341-371 restricted to the per-call pose term; S and tb are hoisted)
//
// Why a CTA pair.  The ablations of the single-CTA kernels (profiles/r1e_harness_ablation.txt) show the first
// layer limited by what one SM can ingest from L2 (~64 B/clk): per 128-edge x 256-output tile it pulls 256 KB of
// gathered pose embeddings, 512 KB of weights and 128 KB of S against 12.3 k tensor-pipe cycles.  Multicasting the
// weights does not help (every SM still receives all of them).  With cta_group::2 the two CTAs of a pair work on
// two consecutive 128-edge tiles (an M = 256 MMA), and each CTA stages only HALF of every weight chunk: the
// tensor cores of both SMs read both halves.  Weight ingest per SM halves, the shared-memory ring stages shrink
// from 48 KB to 32 KB, and that room pays for a 4-deep ring of decoder operand chunks, so the first-layer
// activations H never leave the SM:
//
//   GEMM1  D1[128 x 256] (TMEM cols 0..255)   <- gathered A (cp.async) x W_c,pose chunk (cp.async.bulk)
//   epi-1  16 warps: D1 -> +S +tb -> SiLU -> hi/lo BF16 -> decoder operand chunks in a shared-memory ring
//   GEMM2  D2[128 x 128] (TMEM cols 256.., double buffered) <- those chunks x pose_decoder.0 chunks
//   epi-2  same 16 warps: D2 -> +b -> SiLU -> 128 -> P FMAs -> o
//
// D1 is single-buffered: every epilogue thread pulls its whole share of D1 (64 values) into registers with two TMEM
// loads and releases the accumulator at once (setmaxnreg gives the 16 epilogue warps 104 registers), so GEMM1 of the next
// unit starts ~650 cycles after the commit.  D2 is double-buffered and epilogue-2 is deferred by one unit: GEMM2's MMAs
// queue behind the next GEMM1 on the tensor pipe, and the epilogue never waits for them.
//
//   warps 0-15   epilogues (warp w: TMEM lanes 32 (w & 3).., column group w >> 2)
//   warps 16-19  A gather (cp.async 16 B pieces through the edge index)
//   warp  20     first-layer weight loader (one lane, cp.async.bulk)
//   warp  21     leader CTA: GEMM1 issuer.  peer CTA: relays "my stage is full" to the leader's barrier
//   warp  22     leader CTA: GEMM2 issuer.  peer CTA: relays the decoder-weight ring likewise
//   warp  23     decoder weight loader (one lane)
#pragma once
#include <cuda.h>

#include "kernels_tc.cuh"

namespace ccsp {
namespace tc {

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// The sampling loop is a strict chain node -> edge -> node -> ...; with the launch attribute
// cudaLaunchAttributeProgrammaticStreamSerialization the next kernel's CTAs become resident as soon as SMs free up
// and run their prologue (barrier init, TMEM allocation, weight prefetch) under the tail of the previous kernel.
// pdl_wait() blocks until the previous grid has completed and its writes are visible; every thread calls it before
// its first access to data the previous kernel produced (or still reads).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- TMA (tensor-map) copies in CTA-pair mode --------------------------------------------------------
// With .cta_group::2 the completion may be signalled on the PEER CTA's barrier (given by its shared::cluster
// address), which is what lets both CTAs of a pair feed the leader's "stage full" barrier without a relay.
// 4 rows of a 2-D tensor, picked by index, land as 4 consecutive box-wide rows of the destination tile; the tensor map
// carries the shared-memory swizzle (probed in csrc/tests/tma_gather_test.cu: box {cols, 1}, SWIZZLE_64B reproduces
// sw64_off exactly).
__device__ __forceinline__ void tma_gather4_pair(uint32_t dst, const CUtensorMap *map, int col, int r0, int r1, int r2, int r3,
                                                 uint32_t bar_cluster_addr) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.cta_group::2 "
               "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_tile2d_pair(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar_cluster_addr) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes.cta_group::2 "
               "[%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar_cluster_addr) : "memory");
}

// host: 2-D row-major tensor map (cuTensorMapEncodeTiled fetched through the runtime, libcuda is not linked)
inline cudaError_t make_tmap_2d(CUtensorMap *map, const void *base, uint64_t cols, uint64_t rows, uint64_t row_bytes,
                                uint32_t box_cols, uint32_t box_rows, bool swizzle64) {
  typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiled enc = nullptr;
  if (!enc) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    if (e != cudaSuccess) return e;
    if (!fn) return cudaErrorNotSupported;
    enc = (EncodeTiled)fn;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {row_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// the three tensor maps of the TMA variant of the pair kernel
struct PairMaps {
  CUtensorMap pe;   // pe_split as [n+1 rows][512 BF16] (hi | lo), box {32, 1}, SWIZZLE_64B: gather4 source of the A operand
  CUtensorMap w1;   // first-layer weight blob as rows of 64 B, box {32, 128}: this CTA's half of one operand part
  CUtensorMap wd;   // decoder weight blob as rows of 64 B, box {32, 64}
};
template <class M>
inline cudaError_t make_pair_maps(PairMaps *pm, const void *pe_split, int64_t n_rows, const void *b_blob, int num_types,
                                  const void *w_blob) {
  cudaError_t e = make_tmap_2d(&pm->pe, pe_split, 2 * CCSP_H, (uint64_t)n_rows, M::PE_ROW_BYTES, 32, 1, true);
  if (e != cudaSuccess) return e;
  e = make_tmap_2d(&pm->w1, b_blob, 32, (uint64_t)num_types * 2 * M::NKC1 * M::NS * 256, ROWB, 32, 128, false);
  if (e != cudaSuccess) return e;
  return make_tmap_2d(&pm->wd, w_blob, 32, (uint64_t)M::NKC2 * M::NS * 128, ROWB, 32, 64, false);
}

// ---- cluster / CTA-pair PTX wrappers ---------------------------------------------------------------
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_remote(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
// bulk copy whose complete_tx lands on a barrier given by its shared::cluster address (may be the peer CTA's)
__device__ __forceinline__ void bulk_g2s_rbar(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// The static term S (166 MB) streams through the 126 MB L2 once per evaluation; marking it evict-first keeps the
// hot set (weights 13.6 MB, pose embeddings 9.4 MB) resident, so ring fetches do not pay HBM latency.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void prefetch_l2_bulk(const void *src, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(src), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ float4 ldg_nc_f4_hint(const float4 *p, uint64_t policy) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(policy));
  return v;
}
// (CTA-scope acquire on purpose: a cluster-scope acquire costs an L1 invalidate (CCTL.IVALL) per wait, ~400 cycles in the
// issue loops; what crosses CTAs here are barrier signals, the operand bytes are read by each SM's own tensor core.)
__device__ __forceinline__ bool mbar_try_wait_cl(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t *bar, uint32_t parity) {      // non-blocking
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cl(uint64_t *bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cl(bar, parity)) {
    CCSP_SPIN_GUARD(spins);
  }
}
__device__ __forceinline__ void tmem_alloc2(uint32_t *dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// completion of all MMAs issued so far by this thread -> one arrival on the barrier at this CTA-relative
// offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit2(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// 32 lanes x 16 consecutive FP32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 8 consecutive FP32 columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float *v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// one 64-byte k-chunk of MMAs on the pair: D[256 x N] (+)= A . B^T; every CTA contributes its 128 rows of A and
// its N/2 rows of B, both at the same CTA-relative shared-memory offsets.  b_part = bytes of one operand part
// (hi or lo) of this CTA's B half.
template <class M, int NTILE>
__device__ __forceinline__ void issue_chunk2(uint32_t d_tmem, uint32_t a_hi, uint32_t b_hi, uint32_t b_part, bool first) {
  static_assert(M::KIND == KIND_BF16, "pair kernel is BF16-operand only");
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NTILE >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  const uint32_t a_lo = a_hi + PART, b_lo = b_hi + b_part;
#pragma unroll
  for (int ks = 0; ks < M::KSTEPS; ++ks) {
    const uint32_t ko = ks * 32;
    if (M::NSPLIT == 3) {
      umma2_f16(d_tmem, smem_desc_sw64(a_lo + ko), smem_desc_sw64(b_hi + ko), idesc, !(first && ks == 0));
      umma2_f16(d_tmem, smem_desc_sw64(a_hi + ko), smem_desc_sw64(b_lo + ko), idesc, 1);
      umma2_f16(d_tmem, smem_desc_sw64(a_hi + ko), smem_desc_sw64(b_hi + ko), idesc, 1);
    } else {
      umma2_f16(d_tmem, smem_desc_sw64(a_hi + ko), smem_desc_sw64(b_hi + ko), idesc, !(first && ks == 0));
    }
  }
}

// one k-step (K = 16) of the same
// A128: the A stage holds rows of 128 B, [hi 64 B | lo 64 B] of the k-chunk, in SWIZZLE_128B (the gathered first-layer operand:
// a gathered row lands as ONE 128-byte shared-memory line instead of two 64-byte halves in separate hi / lo tiles)
template <class M, int NTILE, bool A128 = false>
__device__ __forceinline__ void issue_kstep2(uint32_t d_tmem, uint32_t a_hi, uint32_t b_hi, uint32_t b_part, int ks, bool zero) {
  constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NTILE >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  const uint32_t ko = ks * 32;
  const uint32_t b_lo = b_hi + b_part;
  if (A128) {
    static_assert(!A128 || M::NSPLIT == 3, "the 128-byte A rows hold a hi and a lo slice");
    umma2_f16(d_tmem, smem_desc_sw128(a_hi + ROWB + ko), smem_desc_sw64(b_hi + ko), idesc, !zero);
    umma2_f16(d_tmem, smem_desc_sw128(a_hi + ko), smem_desc_sw64(b_lo + ko), idesc, 1);
    umma2_f16(d_tmem, smem_desc_sw128(a_hi + ko), smem_desc_sw64(b_hi + ko), idesc, 1);
    return;
  }
  const uint32_t a_lo = a_hi + PART;
  if (M::NSPLIT == 3) {
    umma2_f16(d_tmem, smem_desc_sw64(a_lo + ko), smem_desc_sw64(b_hi + ko), idesc, !zero);
    umma2_f16(d_tmem, smem_desc_sw64(a_hi + ko), smem_desc_sw64(b_lo + ko), idesc, 1);
    umma2_f16(d_tmem, smem_desc_sw64(a_hi + ko), smem_desc_sw64(b_hi + ko), idesc, 1);
  } else {
    umma2_f16(d_tmem, smem_desc_sw64(a_hi + ko), smem_desc_sw64(b_hi + ko), idesc, !zero);
  }
}

// SiLU with bare ex2/rcp approximations (what __expf/__fdividef use, minus their range fix-ups: 1 + 2^y never
// reaches the rcp overflow range before it is +inf, where x * 0 is the right answer)
__device__ __forceinline__ float silu_raw(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}
// (a, b) -> packed BF16 pair of the rounded values, and of the rounding residuals
__device__ __forceinline__ void split_pair(float a, float b, uint32_t &hi, uint32_t &lo) {
  hi = pack_bf16(a, b);
  lo = pack_bf16(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xFFFF0000u));
}

template <class M>
struct Fused2Cfg {
  static_assert(M::KIND == KIND_BF16, "one 32-column epilogue chunk = one BF16 k-chunk of the decoder");
  static constexpr int NT1 = 256, NT2 = 128;
  // cp.async gather variant with split operands: A rows of ring 1 are 128 B ([hi | lo]) in SWIZZLE_128B (see issue_kstep2)
  static constexpr bool A128 = M::NS == 2;
  static constexpr int B1_PART = (NT1 / 2) * ROWB;            // this CTA's 128 weight rows of one operand part: 8 KB
  static constexpr int B1_STAGE = M::NS * B1_PART;
  static constexpr int B1_BLOB_PART = NT1 * ROWB;             // part size in the packed blob (256 rows)
  static constexpr int B1_BLOB_STAGE = M::NS * B1_BLOB_PART;
  static constexpr int STAGE1 = M::A_STAGE + B1_STAGE;        // 32 KB (x3 split)
  static constexpr int NSTAGE1 = 4;
  static constexpr int A2_STAGE = M::A_STAGE;                 // one decoder k-chunk of 128 rows: 16 KB
  static constexpr int NA2 = 4;                               // one stage per epilogue column group
  static constexpr int W_PART = (NT2 / 2) * ROWB;             // this CTA's 64 decoder-weight rows of one part: 4 KB
  static constexpr int W_STAGE = M::NS * W_PART;
  static constexpr int W_BLOB_PART = NT2 * ROWB;
  static constexpr int W_BLOB_STAGE = M::NS * W_BLOB_PART;
  static constexpr int NW = 2;
  static constexpr int OFF_A2 = NSTAGE1 * STAGE1;
  static constexpr int OFF_W = OFF_A2 + NA2 * A2_STAGE;
  static constexpr int OFF_EXTRA = OFF_W + NW * W_STAGE;
  static constexpr int NUM_EPI = 16, EPI_T = NUM_EPI * 32;
  static constexpr int WARP_PROD0 = NUM_EPI, NUM_PROD_WARPS = 4;
  static constexpr int WARP_LOAD = 20, WARP_MMA1 = 21, WARP_MMA2 = 22, WARP_LOADW = 23;
  static constexpr int THREADS = 24 * 32;                     // 768
  // the pool is what the CTA was launched with (768 x 80): 512 x 104 + 256 x 32 = 61 440 exactly
  static constexpr int EPI_REGS = 104, AUX_REGS = 32;
  // barriers 512 | tb_s 2x256 | bd1 128 | w2t 128x8 | bd2 16 | red 3x128 float4
  static constexpr int SMEM_EXTRA = 512 + (512 + CCSP_HH + CCSP_MAXP * CCSP_HH + 16) * 4 + 3 * SUB_M * 16;
  static constexpr int SMEM_BYTES = OFF_EXTRA + SMEM_EXTRA + 1024;
  static constexpr int D2_COL = 256;
  static_assert(SMEM_BYTES <= 227 * 1024, "pair kernel does not fit in shared memory");
};

// TMA = true: operands arrive through tensor maps (A by tile::gather4, weights as 2-D tiles) whose completion is
// signalled on the leader's barriers directly; TMA = false: cp.async gather + cp.async.bulk + relay lanes.
// PERSIST = true: the kernel loops over all A.num_evals evaluations of a sample() (see FusedArgs): barriers, TMEM and tables
// are set up once, the pipeline counters simply run on, and the kernel boundary is replaced by the node_done / edge_done flags.
template <class M, bool TMA, bool PERSIST = false>
__global__ void __launch_bounds__(Fused2Cfg<M>::THREADS, 1)
k_edge_fused2_tc(const FusedArgs A, const __grid_constant__ PairMaps maps) {
  static_assert(!(TMA && PERSIST), "the persistent variant uses the cp.async gather");
  using C = Fused2Cfg<M>;
  // Persistent mode walks "chain evaluations" q = 0, 1, ..: evaluation q / n_ch of chain q % n_ch (FusedArgs::num_chains).
  // Units of successive chain evaluations are dealt round-robin to the CTA pairs as ONE stream (the pair that took the last
  // unit of q is followed by its neighbour for the first unit of q + 1), so no pair idles on a ragged last round.
  const int n_ch = (PERSIST && A.num_chains > 1) ? A.num_chains : 1;
  const int n_q = PERSIST ? A.num_evals * n_ch : 1;
#define CHAIN_EVAL(q)                                                                                       \
  const int ch = (PERSIST && n_ch > 1) ? (q) % n_ch : 0, ev = PERSIST ? (q) / n_ch : 0;                     \
  const int tile0 = (PERSIST && n_ch > 1) ? A.chain_tile0[ch] : 0;                                          \
  const int num_units = (PERSIST && n_ch > 1) ? A.chain_tile0[ch + 1] - tile0 : num_units_all;              \
  int ufirst = unit0 - cbase;                                                                               \
  if (ufirst < 0) ufirst += unit_step;                                                                      \
  if (PERSIST) cbase = (cbase + num_units) % unit_step;                                                     \
  (void)ev; (void)ch
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *extra = smem + C::OFF_EXTRA;
  uint64_t *full1 = reinterpret_cast<uint64_t *>(extra);     // [4] stage of ring 1 filled (leader: incl. the peer's)
  uint64_t *empty1 = full1 + 4;                              // [4]
  uint64_t *a2_full = empty1 + 4;                            // [4] decoder operand chunk written (leader only)
  uint64_t *a2_empty = a2_full + 4;                          // [4][2] per 32-byte half (k-step) of a group's stage
  uint64_t *w_full = a2_empty + 8;                           // [2]
  uint64_t *w_empty = w_full + 2;                            // [2]
  uint64_t *tfull1 = w_empty + 2, *tempty1 = tfull1 + 1;     // D1
  uint64_t *tfull2 = tempty1 + 1, *tempty2 = tfull2 + 2;     // D2 [2]
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tempty2 + 2);
  float *tb_s = reinterpret_cast<float *>(extra + 512);      // [2][256]
  float *bd1 = tb_s + 512;                                   // [128]
  float *w2t = bd1 + CCSP_HH;                                // [128][8]
  float *bd2 = w2t + CCSP_MAXP * CCSP_HH;                    // [8] (+8 pad)
  float4 *red = reinterpret_cast<float4 *>(bd2 + 16);        // [3][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_units_all = (A.num_m_tiles / 2) * 2;         // (pair of 128-edge tiles, slot)
  const int unit0 = blockIdx.x >> 1, unit_step = gridDim.x >> 1;
  const uint32_t smem_base = smem_u32(smem);
  // (Letting the peer's NON-tensor cp.async.bulk complete_tx on the leader's barrier traps on B200: its barrier has to
  // live in the destination CTA.  Hence the relay lanes of the TMA = false variant.)
  long long *const tr = (!PERSIST && A.trace && blockIdx.x == 0) ? A.trace : nullptr;
#define TR(role, slot) do { if (tr && it < 8) tr[((role) * 8 + it) * 16 + (slot)] = clock64(); } while (0)
#define TRP(role, slot) do { if (tr && itp < 8) tr[((role) * 8 + itp) * 16 + (slot)] = clock64(); } while (0)

  const long long t_entry = tr ? clock64() : 0;
  if (threadIdx.x == 0) pdl_launch_dependents();
  if (PERSIST && threadIdx.x == 0 && A.arrive) {
    atomicAdd_system(A.arrive, 1u);
    __threadfence_system();
  }
  if (threadIdx.x == 0) {
    // ring 1: own gather threads + own weight loader (+ at the leader: the peer's relay lane)
    // (TMA variant: leader only, one arrive.expect_tx per loader lane of the pair: 2 x A + 2 x weights)
    for (int s = 0; s < C::NSTAGE1; ++s) { mbar_init(&full1[s], TMA ? 4 : C::NUM_PROD_WARPS * 32 + 1 + (leader ? 1 : 0)); mbar_init(&empty1[s], 1); }
    for (int s = 0; s < C::NA2; ++s) { mbar_init(&a2_full[s], 8); mbar_init(&a2_empty[2 * s], 1); mbar_init(&a2_empty[2 * s + 1], 1); }
    for (int s = 0; s < C::NW; ++s) { mbar_init(&w_full[s], (TMA || leader) ? 2 : 1); mbar_init(&w_empty[s], 1); }
    mbar_init(tfull1, 1); mbar_init(tempty1, 2 * C::NUM_EPI);
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull2[b], 1); mbar_init(&tempty2[b], 2 * C::NUM_EPI); }
    fence_barrier_init();
  }
  if (warp == C::WARP_MMA1) tmem_alloc2(tmem_ptr, 512);
  if (warp < C::NUM_EPI) {
    for (int i = threadIdx.x; i < CCSP_HH; i += C::EPI_T) bd1[i] = A.bd1[i];
    for (int i = threadIdx.x; i < CCSP_MAXP * CCSP_HH; i += C::EPI_T) {
      const int j = i / CCSP_MAXP, pp = i % CCSP_MAXP;
      w2t[i] = pp < A.P ? A.Wd2[pp * CCSP_HH + j] : 0.f;
    }
    if (threadIdx.x < CCSP_MAXP) bd2[threadIdx.x] = threadIdx.x < A.P ? A.bd2[threadIdx.x] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();                      // both CTAs' barriers are initialised before any remote arrival / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // Register re-partition (warpgroup-granular): the 16 epilogue warps take 104 registers so that a thread can hold
  // its whole share of D1 (64 values) and release the accumulator right after two TMEM loads; the copy / issue /
  // relay warps need few.
#define REG_DEC() asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::AUX_REGS))
#define REG_INC() asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::EPI_REGS))

  if (warp >= C::WARP_PROD0 && warp < C::WARP_PROD0 + C::NUM_PROD_WARPS) {
    REG_DEC();
    if (TMA) {
      // ============ A by TMA gather4: one warp, lane l owns rows 4l .. 4l+3 of this CTA's tile ===========
      if (warp == C::WARP_PROD0) {
        uint32_t g = 0;
        uint32_t rfull[C::NSTAGE1];
#pragma unroll
        for (int s = 0; s < C::NSTAGE1; ++s) rfull[s] = mapa_u32(smem_u32(&full1[s]), 0);
        pdl_wait();                      // pe_split is written by the preceding node kernel
        for (int u = unit0; u < num_units_all; u += unit_step) {
          const int m0 = ((u >> 1) * 2 + (int)rank) * SUB_M;
          int4 idx = make_int4(0, 0, 0, 0);
#pragma unroll 1
          for (int kc = 0; kc < M::NKC1; ++kc, ++g) {
            if (kc == 0 || kc == M::NKC1 / 2)
              idx = __ldg(reinterpret_cast<const int4 *>((kc == 0 ? A.src_i : A.src_j) + m0) + lane);
            const uint32_t s = g % C::NSTAGE1;
            mbar_wait(&empty1[s], ((g / C::NSTAGE1) & 1) ^ 1);
            if (A.dbg & 1) { if (lane == 0) mbar_arrive_remote(rfull[s]); continue; }
            if (lane == 0) mbar_arrive_expect_tx_remote(rfull[s], M::A_STAGE);
            const uint32_t dst = smem_base + s * C::STAGE1 + lane * 4 * ROWB;
            const int col = M::pe_off(kc % (M::NKC1 / 2), 0, 0) / M::ELT;       // element column of the hi slice; lo = + KC
            tma_gather4_pair(dst, &maps.pe, col, idx.x, idx.y, idx.z, idx.w, rfull[s]);
            if (M::NS == 2) tma_gather4_pair(dst + PART, &maps.pe, col + M::KC, idx.x, idx.y, idx.z, idx.w, rfull[s]);
          }
        }
      }
    } else {
    // ============ A gather: this CTA's 128 edges ========================================================
    // x3 split: thread = (piece q8 of the row's 128 contiguous bytes [hi 64 | lo 64] of this k-chunk, rows r0 + 16 p);
    // a warp instruction covers 4 rows x 128 B = 4 cache lines / 4 shared-memory wavefronts.  Single pass: hi only.
    const int t = threadIdx.x - C::WARP_PROD0 * 32;
    constexpr int LPR = M::NS == 2 ? 8 : 4;        // lanes per row
    constexpr int RPP = 128 / LPR;                 // rows per pass
    constexpr int NP = SUB_M / RPP;                // passes (= cp.async per thread and chunk)
    const int q8 = t % LPR, r0 = t / LPR;
    const int part = q8 >> 2, q = q8 & 3;
    // Completion is signalled by the copy engine itself (cp.async.mbarrier.arrive.noinc: one arrival per thread
    // when all its earlier cp.async have landed), so a thread never waits for data and all NSTAGE1 stages can be
    // in flight; the peer's barrier is forwarded to the leader by the relay lane below.
    uint32_t g = 0, it = 0;
    uint32_t roff[NP];                   // row offsets in 16-byte units (row stride 1 KB: fits 32 bits up to 4 M nodes)
    uint32_t live = 0;                   // bit p: row p of this thread is a real edge
    auto load_rows = [&](const int *idx, int m0) {
      live = 0;
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const int ix = __ldg(&idx[m0 + r0 + RPP * p]);
        roff[p] = (uint32_t)ix * (M::PE_ROW_BYTES / 16);
        live |= (ix + 1 != A.pad_row_plus1 ? 1u : 0u) << p;      // padded rows are zero-filled without touching memory
      }
    };
    // the first unit's edge indices are plan constants: fetched under the tail of the preceding node kernel
    bool preloaded = false;
    if (!PERSIST && unit0 < num_units_all) {
      load_rows(A.src_i, ((unit0 >> 1) * 2 + (int)rank) * SUB_M);
      preloaded = true;
    }
    if (!PERSIST) pdl_wait();            // pe_split is written by the preceding node kernel
    int cbase = 0;
    for (int qe = 0; qe < n_q; ++qe) {   // (q is this thread's piece index)
    CHAIN_EVAL(qe);
    if (PERSIST) {                       // ... or by iteration ev of the persistent node kernel (one poller per CTA)
      if (n_ch == 1 || ufirst < num_units) {      // (with chains: nothing to gather here, nothing to wait for)
        if (t == 0) {
          wait_flag_ge(A.node_done + 32 * ch, (unsigned)(ev + 1) * (n_ch > 1 ? A.chain_nblk[ch] : A.node_ctas));
          if (blockIdx.x == 0) PTRACE(A.trace, 4, qe);
        }
        asm volatile("bar.sync 3, 128;" ::: "memory");
      }
    }
    for (int u = ufirst; u < num_units; u += unit_step, ++it) {
      const int m0 = (tile0 + (u >> 1) * 2 + (int)rank) * SUB_M;
#pragma unroll 1
      for (int kc = 0; kc < M::NKC1; ++kc, ++g) {
        if (kc == 0 || kc == M::NKC1 / 2) {
          if (kc == 0 && preloaded) preloaded = false;
          else load_rows(kc == 0 ? A.src_i : A.src_j, m0);
        }
        const uint32_t s = g % C::NSTAGE1;
        mbar_wait(&empty1[s], ((g / C::NSTAGE1) & 1) ^ 1);
        if (t == 0) TR(6, kc);
        if (!(A.dbg & 1)) {
          const uint32_t koff = M::pe_off(kc % (M::NKC1 / 2), part, q);
          const uint32_t st = smem_base + s * C::STAGE1 + (C::A128 ? 0 : part * PART);
#pragma unroll
          for (int p = 0; p < NP; ++p)
            cp_async16_zfill(st + (C::A128 ? sw128_off(r0 + RPP * p, q8) : sw64_off(r0 + RPP * p, q)),
                             A.pe_split + (size_t)roff[p] * 16 + koff, ((live >> p) & 1u) ? 16u : 0u);
        }
        cp_async_arrive_noinc(&full1[s]);
        if (t == 0) TR(7, kc);
      }
    }
    }
    }
  } else if (warp >= C::WARP_LOAD) {
    REG_DEC();       // one instruction for the whole warpgroup (warps 20-23), then the per-warp roles
    if (warp == C::WARP_LOAD) {
    if (lane == 0) {
      // ============ first-layer weights: this CTA's 128 of the 256 rows of every chunk =================
      uint32_t g = 0;
      int cbase = 0;
      for (int q = 0; q < n_q; ++q) {
      CHAIN_EVAL(q);
      for (int u = ufirst; u < num_units; u += unit_step) {
        const int mt = tile0 + (u >> 1) * 2, slot = u & 1;
        const int grp = __ldg(&A.tile_type[mt]);
        const uint8_t *blob = A.b_blob + ((size_t)(grp * 2 + slot) * M::NKC1) * C::B1_BLOB_STAGE + rank * C::B1_PART;
        for (int kc = 0; kc < M::NKC1; ++kc, ++g) {
          const uint32_t s = g % C::NSTAGE1;
          mbar_wait(&empty1[s], ((g / C::NSTAGE1) & 1) ^ 1);
          // TMA variant: the copy's complete_tx lands on the LEADER's barrier; relay variant: on the local one
          const uint32_t bar = TMA ? mapa_u32(smem_u32(&full1[s]), 0) : smem_u32(&full1[s]);
          if (A.dbg & 2) { mbar_arrive_remote(bar); continue; }
          mbar_arrive_expect_tx_remote(bar, C::B1_STAGE);
          const uint32_t dst = smem_base + s * C::STAGE1 + M::A_STAGE;
          if (TMA) {
            const int row = ((grp * 2 + slot) * M::NKC1 + kc) * (M::NS * C::NT1) + (int)rank * (C::NT1 / 2);
            tma_tile2d_pair(dst, &maps.w1, 0, row, bar);
            if (M::NS == 2) tma_tile2d_pair(dst + C::B1_PART, &maps.w1, 0, row + C::NT1, bar);
            continue;
          }
          const uint8_t *src = blob + (size_t)kc * C::B1_BLOB_STAGE;
          bulk_g2s_rbar(dst, src, C::B1_PART, bar);
          if (M::NS == 2) bulk_g2s_rbar(dst + C::B1_PART, src + C::B1_BLOB_PART, C::B1_PART, bar);
        }
      }
      }
    }
  } else if (warp == C::WARP_LOADW) {
    // (own warp: a lane that sleeps in mbarrier.try_wait holds up the other lanes of its warp)
    if (lane == 0) {
      // ============ decoder weights: this CTA's 64 of the 128 rows, chunks in GEMM2's consumption order ==
      uint32_t g2 = 0;
      const uint8_t *blob = A.w_blob + rank * C::W_PART;
      int cbase = 0;
      for (int qe = 0; qe < n_q; ++qe) {
      CHAIN_EVAL(qe);
      for (int u = ufirst; u < num_units; u += unit_step) {
        for (int q = 0; q < 8; ++q, ++g2) {
          const uint32_t s = g2 % C::NW;
          mbar_wait(&w_empty[s], ((g2 / C::NW) & 1) ^ 1);
          const uint32_t bar = TMA ? mapa_u32(smem_u32(&w_full[s]), 0) : smem_u32(&w_full[s]);
          if (A.dbg & 2) { mbar_arrive_remote(bar); continue; }
          const int c = 2 * (q & 3) + (q >> 2);
          mbar_arrive_expect_tx_remote(bar, C::W_STAGE);
          const uint32_t dst = smem_base + C::OFF_W + s * C::W_STAGE;
          if (TMA) {
            const int row = c * (M::NS * C::NT2) + (int)rank * (C::NT2 / 2);
            tma_tile2d_pair(dst, &maps.wd, 0, row, bar);
            if (M::NS == 2) tma_tile2d_pair(dst + C::W_PART, &maps.wd, 0, row + C::NT2, bar);
            continue;
          }
          const uint8_t *src = blob + (size_t)c * C::W_BLOB_STAGE;
          bulk_g2s_rbar(dst, src, C::W_PART, bar);
          if (M::NS == 2) bulk_g2s_rbar(dst + C::W_PART, src + C::W_BLOB_PART, C::W_PART, bar);
        }
      }
      }
    }
  } else if (warp == C::WARP_MMA1) {
    if (lane == 0) {
      if (leader) {
        // ============ GEMM1 issuer (M = 256 over the pair) =============================================
        uint32_t g = 0, it = 0;
        if (tr) tr[15] = t_entry;          // kernel entry of this thread (slot 15 of MMA1 / unit 0)
        int cbase = 0;
        for (int q = 0; q < n_q; ++q) {
        CHAIN_EVAL(q);
        for (int u = ufirst; u < num_units; u += unit_step, ++it) {
          TR(0, 0);
          mbar_wait_cl(tempty1, (it & 1) ^ 1);
          TR(0, 1);
          tc_fence_after();
          bool ready = false;              // the stage of this chunk was already seen full by the peek below
          for (int kc = 0; kc < M::NKC1; ++kc, ++g) {
            const uint32_t s = g % C::NSTAGE1;
            if (!ready) mbar_wait_cl(&full1[s], (g / C::NSTAGE1) & 1);
            if (PERSIST && blockIdx.x == 0 && ev == 5) PTRACE(A.trace, 8, kc);
            if (kc == 0) TR(0, 2);
            if (kc == 8) TR(0, 3);
            tc_fence_after();
            const uint32_t a_hi = smem_base + s * C::STAGE1;
            if (!(A.dbg & 8)) issue_kstep2<M, C::NT1, C::A128 && !TMA>(tmem_base, a_hi, a_hi + M::A_STAGE, C::B1_PART, 0, kc == 0);
            // peek at the next stage while this chunk's MMAs are queued, so that its first MMA can follow the commit
            // without the barrier round trip (non-blocking: a blocking wait here would delay this stage's release)
            ready = kc + 1 < M::NKC1 && mbar_test_wait(&full1[(g + 1) % C::NSTAGE1], ((g + 1) / C::NSTAGE1) & 1);
            if (!(A.dbg & 8)) issue_kstep2<M, C::NT1, C::A128 && !TMA>(tmem_base, a_hi, a_hi + M::A_STAGE, C::B1_PART, 1, false);
            umma_commit2(&empty1[s]);
          }
          umma_commit2(tfull1);
          if (PERSIST && blockIdx.x == 0) PTRACE(A.trace, 5, (int)it);
          TR(0, 4);
        }
        }
      } else if (!TMA) {
        // ============ peer: forward "stage s is full here (A rows + weight half)" to the leader ==========
        uint32_t g = 0;
        uint32_t rfull[C::NSTAGE1];
#pragma unroll
        for (int s = 0; s < C::NSTAGE1; ++s) rfull[s] = mapa_u32(smem_u32(&full1[s]), 0);
        int cbase = 0;
        for (int q = 0; q < n_q; ++q) {
        CHAIN_EVAL(q);
        for (int u = ufirst; u < num_units; u += unit_step) {
          for (int kc = 0; kc < M::NKC1; ++kc, ++g) {
            const uint32_t s = g % C::NSTAGE1;
            mbar_wait(&full1[s], (g / C::NSTAGE1) & 1);
            mbar_arrive_remote(rfull[s]);
          }
        }
        }
      }
    }
  } else if (warp == C::WARP_MMA2) {
    if (lane == 0) {
      if (leader) {
        // ============ GEMM2 issuer ======================================================================
        uint32_t g2 = 0, it = 0;
        int cbase = 0;
        for (int qe = 0; qe < n_q; ++qe) {
        CHAIN_EVAL(qe);
        for (int u = ufirst; u < num_units; u += unit_step, ++it) {
          const uint32_t buf = it & 1;
          TR(1, 0);
          mbar_wait_cl(&tempty2[buf], ((it >> 1) & 1) ^ 1);
          TR(1, 1);
          tc_fence_after();
          for (int q = 0; q < 8; ++q, ++g2) {
            const uint32_t cg = q & 3, uu = it * 2 + (q >> 2), ws = g2 % C::NW;
            mbar_wait_cl(&a2_full[cg], uu & 1);
            if (q == 0) TR(1, 2);
            if (q == 4) TR(1, 4);
            mbar_wait_cl(&w_full[ws], (g2 / C::NW) & 1);
            if (q == 0) TR(1, 3);
            if (q == 4) TR(1, 5);
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {       // each 32-byte half of the group's stage is released on its own
              if (!(A.dbg & 8))
                issue_kstep2<M, C::NT2>(tmem_base + C::D2_COL + buf * C::NT2, smem_base + C::OFF_A2 + cg * C::A2_STAGE,
                                        smem_base + C::OFF_W + ws * C::W_STAGE, C::W_PART, ks, q == 0 && ks == 0);
              umma_commit2(&a2_empty[2 * cg + ks]);
            }
            umma_commit2(&w_empty[ws]);
          }
          umma_commit2(&tfull2[buf]);
          TR(1, 6);
        }
        }
      } else if (!TMA) {
        uint32_t g2 = 0;
        uint32_t rfull[C::NW];
#pragma unroll
        for (int s = 0; s < C::NW; ++s) rfull[s] = mapa_u32(smem_u32(&w_full[s]), 0);
        int cbase = 0;
        for (int qe = 0; qe < n_q; ++qe) {
        CHAIN_EVAL(qe);
        for (int u = ufirst; u < num_units; u += unit_step) {
          for (int q = 0; q < 8; ++q, ++g2) {
            const uint32_t s = g2 % C::NW;
            mbar_wait(&w_full[s], (g2 / C::NW) & 1);
            mbar_arrive_remote(rfull[s]);
          }
        }
        }
      }
    }
  }
  } else if (warp < C::NUM_EPI) {
    REG_INC();
    // ============ epilogues: warp w <-> TMEM lanes 32 (w & 3).., column group cg = w >> 2 =================
    const int quarter = warp & 3, cg = warp >> 2;
    const int r = quarter * 32 + lane;
    uint8_t *a2_stage = smem + C::OFF_A2 + cg * C::A2_STAGE;
    const uint32_t r_tempty1 = mapa_u32(smem_u32(tempty1), 0);
    const uint32_t r_tempty2_0 = mapa_u32(smem_u32(&tempty2[0]), 0), r_tempty2_1 = mapa_u32(smem_u32(&tempty2[1]), 0);
    const uint32_t r_a2_full = mapa_u32(smem_u32(&a2_full[cg]), 0);
    const uint64_t pol_s = l2_policy_evict_first();
    if (threadIdx.x < 4 && unit0 < num_units_all && n_ch == 1 && !(A.dbg & 4)) {
      // first unit's slice of S -> L2 while the ring fills (S is a plan constant: no need to wait for the previous kernel)
      const size_t rb = (size_t)((unit0 >> 1) * 2 + (int)rank) * 4 + threadIdx.x;
      prefetch_l2_bulk(A.S + (rb * 16 + (size_t)(unit0 & 1) * 8) * 1024, 32768, pol_s);
    }
    // ---- epilogue-2 of unit itp (deferred by one unit: GEMM2 had a whole GEMM1 to finish): D2 -> o -------
    auto epi2 = [&](uint32_t itp, size_t rowp, int slotp) {
      if (!PERSIST && itp == 0) pdl_wait();   // o is still being read by the preceding node kernel until it completes
                                              // (persistent: the gathers of this evaluation already waited for the node iteration)
      const int trole = warp == 0 ? 2 : 3;
      const bool tron = lane == 0 && (warp == 0 || warp == 12);
      const uint32_t buf = itp & 1;
      if (tron) TRP(trole, 7);
      mbar_wait_cl(&tfull2[buf], (itp >> 1) & 1);
      if (PERSIST && blockIdx.x == 0 && threadIdx.x == 0) PTRACE(A.trace, 9, 3 + (int)(itp & 3));
      if (tron) TRP(trole, 8);
      tc_fence_after();
      const uint32_t taddr2 = tmem_base + C::D2_COL + buf * C::NT2 + cg * 32 + ((uint32_t)(quarter * 32) << 16);
      float acc[CCSP_MAXP];
#pragma unroll
      for (int p = 0; p < CCSP_MAXP; ++p) acc[p] = 0.f;
#pragma unroll 1
      for (int hc = 0; hc < 2; ++hc) {
        float v[16];
        tmem_ld16(taddr2 + hc * 16, v);
        if (hc == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(buf ? r_tempty2_1 : r_tempty2_0);
        }
        const int c0 = cg * 32 + hc * 16;
        if (A.P <= 4) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float d = silu_raw(v[j] + bd1[c0 + j]);
            const float4 w = *reinterpret_cast<const float4 *>(&w2t[(c0 + j) * CCSP_MAXP]);
            acc[0] = fmaf(d, w.x, acc[0]); acc[1] = fmaf(d, w.y, acc[1]); acc[2] = fmaf(d, w.z, acc[2]); acc[3] = fmaf(d, w.w, acc[3]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float d = silu_raw(v[j] + bd1[c0 + j]);
            const float4 w0 = *reinterpret_cast<const float4 *>(&w2t[(c0 + j) * CCSP_MAXP]);
            const float4 w1 = *reinterpret_cast<const float4 *>(&w2t[(c0 + j) * CCSP_MAXP + 4]);
            acc[0] = fmaf(d, w0.x, acc[0]); acc[1] = fmaf(d, w0.y, acc[1]); acc[2] = fmaf(d, w0.z, acc[2]); acc[3] = fmaf(d, w0.w, acc[3]);
            acc[4] = fmaf(d, w1.x, acc[4]); acc[5] = fmaf(d, w1.y, acc[5]); acc[6] = fmaf(d, w1.z, acc[6]); acc[7] = fmaf(d, w1.w, acc[7]);
          }
        }
      }
      if (tron) TRP(trole, 9);
      // column groups 1..3 hand their partial sums to group 0 (fixed order -> deterministic)
      float *orow = A.o + ((size_t)rowp * 2 + slotp) * A.P;
      if (cg > 0) red[(cg - 1) * SUB_M + r] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      asm volatile("bar.sync 2, 512;" ::: "memory");
      if (cg == 0 && !(A.dbg & 4)) {
        const float4 r1 = red[r], r2 = red[SUB_M + r], r3 = red[2 * SUB_M + r];
        float res[4] = {((acc[0] + r1.x) + r2.x) + r3.x + bd2[0], ((acc[1] + r1.y) + r2.y) + r3.y + bd2[1],
                        ((acc[2] + r1.z) + r2.z) + r3.z + bd2[2], ((acc[3] + r1.w) + r2.w) + r3.w + bd2[3]};
        if (A.P == 4) {
          *reinterpret_cast<float4 *>(orow) = make_float4(res[0], res[1], res[2], res[3]);
        } else {
#pragma unroll
          for (int p = 0; p < 4; ++p)
            if (p < A.P) orow[p] = res[p];
        }
      }
      if (tron) TRP(trole, 10);
      if (A.P > 4) {                                // second round for components 4..7 (robot poses, P = 5)
        asm volatile("bar.sync 2, 512;" ::: "memory");
        if (cg > 0) red[(cg - 1) * SUB_M + r] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        asm volatile("bar.sync 2, 512;" ::: "memory");
        if (cg == 0 && !(A.dbg & 4)) {
          const float4 r1 = red[r], r2 = red[SUB_M + r], r3 = red[2 * SUB_M + r];
          float res[4] = {((acc[4] + r1.x) + r2.x) + r3.x + bd2[4], ((acc[5] + r1.y) + r2.y) + r3.y + bd2[5],
                          ((acc[6] + r1.z) + r2.z) + r3.z + bd2[6], ((acc[7] + r1.w) + r2.w) + r3.w + bd2[7]};
#pragma unroll
          for (int p = 0; p < 4; ++p)
            if (4 + p < A.P) orow[4 + p] = res[p];
        }
      }
    };
    size_t prev_row = 0;
    int prev_slot = 0;
    uint32_t it = 0;
    int cbase = 0;
    // Epilogue-2 is deferred by one unit.  Launch-per-evaluation mode drains it at the end of the launch; the persistent mode
    // carries it ACROSS chain evaluations (the tensor pipe already runs the next chain's first GEMM1, and GEMM2 of the pending
    // unit queues behind it): `o` of chain evaluation q is then complete — and edge_done signalled — after epilogue-1 of the
    // first unit of q + 1.  It drains only when this pair has no unit in q + 1 or the sample ends.
    bool pend = false, pend_last = false;
    int pend_ch = 0, pend_q = 0;
    unsigned pend_cnt = 1;
    auto signal_done = [&](int c, int qq, unsigned cnt) {   // every o row of this CTA for that chain evaluation is written: tell the node kernel
      __threadfence();
      asm volatile("bar.sync 2, 512;" ::: "memory");
      if (threadIdx.x == 0) {
        red_release_gpu_add(A.edge_done + 32 * c, cnt);
        if (blockIdx.x == 0) PTRACE(A.trace, 7, qq);
      }
    };
    for (int q = 0; q < n_q; ++q) {
    CHAIN_EVAL(q);
    const float *tb_ev = PERSIST ? A.tb_base + (size_t)__ldg(&A.eval_t[ev]) * A.tb_stride : A.tb;
    for (int u = ufirst; u < num_units; u += unit_step, ++it) {
      const int mt = tile0 + (u >> 1) * 2 + (int)rank, slot = u & 1;
      const int grp = __ldg(&A.tile_type[mt]);
      const size_t row = (size_t)mt * SUB_M + r;
      const int gcol0 = slot * 256 + cg * 64;
      // S in the blocked layout: 32-row x 32-col blocks of 8 pieces x 32 lanes x 16 B (coalesced 512 B / instruction)
      const float4 *Sblk = reinterpret_cast<const float4 *>(A.S) + (((row >> 5) * 16 + (gcol0 >> 5)) * 8) * 32 + lane;
      const bool noS = (A.dbg & 4) != 0;
      float4 sa = noS ? make_float4(0.f, 0.f, 0.f, 0.f) : ldg_nc_f4_hint(Sblk, pol_s), sb = noS ? sa : ldg_nc_f4_hint(Sblk + 32, pol_s);
      if (threadIdx.x < 4 && u + unit_step < num_units && !(A.dbg & 4)) {
        // next unit's slice of S (4 row blocks x 32 KB contiguous) -> L2, so the register prefetch below only
        // has to cover an L2 hit
        const int un = u + unit_step;
        const size_t rb = (size_t)(tile0 + (un >> 1) * 2 + (int)rank) * 4 + threadIdx.x;
        prefetch_l2_bulk(A.S + (rb * 16 + (size_t)(un & 1) * 8) * 1024, 32768, pol_s);
      }
      float *tbu = tb_s + (it & 1) * 256;
      if (threadIdx.x < 256) tbu[threadIdx.x] = __ldg(&tb_ev[(size_t)grp * CCSP_H2 + slot * 256 + threadIdx.x]);
      const int trole = warp == 0 ? 2 : 3;
      const bool tron = lane == 0 && (warp == 0 || warp == 12);
      if (tron) TR(trole, 0);
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (tron) TR(trole, 1);
      // ---- epilogue-1: D1 -> decoder operand chunks -------------------------------------------------
      mbar_wait_cl(tfull1, it & 1);
      if (PERSIST && blockIdx.x == 0 && threadIdx.x == 0) PTRACE(A.trace, 6, q);
      if (tron) TR(trole, 2);
      tc_fence_after();
      const uint32_t taddr1 = tmem_base + cg * 64 + ((uint32_t)(quarter * 32) << 16);
      float vall[64];
      tmem_ld32(taddr1, vall);
      tmem_ld32(taddr1 + 32, vall + 32);
      tc_fence_before();                             // D1 fully in registers: GEMM1 of the next unit may overwrite it
      if (PERSIST && blockIdx.x == 0 && threadIdx.x == 0 && ev == 5) PTRACE(A.trace, 9, 0);
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(r_tempty1);
#pragma unroll
      for (int half = 0; half < 2; ++half) {         // 32 columns = one decoder k-chunk
        const float *v = vall + half * 32;
        if (tron) TR(trole, 3 + 2 * half);
#pragma unroll
        for (int pc = 0; pc < 4; ++pc) {             // 8 columns -> one 16-byte piece of hi and of lo
          const int pi = half * 4 + pc;
          const float4 ca = sa, cb = sb;
          if (pi < 7 && !noS) {                      // S of the next piece, one step ahead
            const float4 *nx = Sblk + ((pi + 1) >> 2) * 256 + ((pi + 1) & 3) * 64;
            sa = ldg_nc_f4_hint(nx, pol_s); sb = ldg_nc_f4_hint(nx + 32, pol_s);
          }
          const float4 ta = *reinterpret_cast<const float4 *>(&tbu[cg * 64 + pi * 8]);
          const float4 tb4 = *reinterpret_cast<const float4 *>(&tbu[cg * 64 + pi * 8 + 4]);
          float f[8];
          f[0] = silu_raw(v[pc * 8 + 0] + ca.x + ta.x); f[1] = silu_raw(v[pc * 8 + 1] + ca.y + ta.y);
          f[2] = silu_raw(v[pc * 8 + 2] + ca.z + ta.z); f[3] = silu_raw(v[pc * 8 + 3] + ca.w + ta.w);
          f[4] = silu_raw(v[pc * 8 + 4] + cb.x + tb4.x); f[5] = silu_raw(v[pc * 8 + 5] + cb.y + tb4.y);
          f[6] = silu_raw(v[pc * 8 + 6] + cb.z + tb4.z); f[7] = silu_raw(v[pc * 8 + 7] + cb.w + tb4.w);
          uint4 hi, lo;
          split_pair(f[0], f[1], hi.x, lo.x); split_pair(f[2], f[3], hi.y, lo.y);
          split_pair(f[4], f[5], hi.z, lo.z); split_pair(f[6], f[7], hi.w, lo.w);
          // GEMM2 is done with this 32-byte half (k-step) of the group's stage
          if ((pc & 1) == 0) mbar_wait(&a2_empty[2 * cg + (pc >> 1)], ((it * 2 + half) & 1) ^ 1);
          uint8_t *dst = a2_stage + sw64_off(r, pc);
          *reinterpret_cast<uint4 *>(dst) = hi;
          if (M::NS == 2) *reinterpret_cast<uint4 *>(dst + PART) = lo;
        }
        fence_proxy_async();                         // chunk complete: publish to the async proxy, tell the leader
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(r_a2_full);
        if (PERSIST && blockIdx.x == 0 && threadIdx.x == 0 && ev == 5) PTRACE(A.trace, 9, 1 + half);
        if (tron) TR(trole, 4 + 2 * half);
      }
      if (pend) {
        epi2(it - 1, prev_row, prev_slot);
        if (PERSIST && pend_last) signal_done(pend_ch, pend_q, pend_cnt);
      }
      pend = true; pend_last = u + unit_step >= num_units; pend_ch = ch; pend_q = q;
      // one count per CTA and evaluation (single chain) or per unit of the chain evaluation (chains)
      pend_cnt = n_ch > 1 ? (unsigned)((num_units - ufirst + unit_step - 1) / unit_step) : 1u;
      prev_row = row; prev_slot = slot;
    }
    if (PERSIST && n_ch == 1 && ufirst >= num_units) signal_done(ch, q, 1u);   // (single chain: every CTA reports every evaluation)
    bool drain = pend;
    if (PERSIST && pend && q + 1 < n_q) {                       // keep it pending if a unit of q + 1 follows on this pair
      const int cn = n_ch > 1 ? (q + 1) % n_ch : 0;
      const int units_n = n_ch > 1 ? A.chain_tile0[cn + 1] - A.chain_tile0[cn] : num_units_all;
      int uf = unit0 - cbase;
      if (uf < 0) uf += unit_step;
      drain = uf >= units_n || A.drain_each_eval;
    }
    if (drain) {
      epi2(it - 1, prev_row, prev_slot);
      if (PERSIST) signal_done(pend_ch, pend_q, pend_cnt);
      pend = false;
    }
    }
  }
#undef TR
#undef TRP
#undef CHAIN_EVAL
#undef REG_DEC
#undef REG_INC
  tc_fence_before();
  __syncthreads();
  cluster_sync();                      // no CTA exits (or frees TMEM) while its peer may still signal it or read its operands
  if (warp == C::WARP_MMA1) tmem_dealloc2(tmem_base, 512);
}

template <class M, bool TMA>
cudaError_t launch_fused2_impl(const FusedArgs &a, const PairMaps &maps, int num_sms, cudaStream_t st) {
  if (a.num_evals != 0) return cudaErrorInvalidValue;      // persistent arguments go through launch_fused2_persistent
  using C = Fused2Cfg<M>;
  static int max_clusters_dev[64] = {};      // per device: function attributes and cluster occupancy
  int dev_ = 0;
  cudaGetDevice(&dev_);
  int &max_clusters = max_clusters_dev[dev_ & 63];
  if (max_clusters == 0) {
    cudaError_t e = cudaFuncSetAttribute(k_edge_fused2_tc<M, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    // setmaxnreg moves registers inside the pool the CTA was launched with; if the compiler ever allocated fewer than
    // the re-partition needs, the epilogue warps would wait for registers forever — refuse to launch instead
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, k_edge_fused2_tc<M, TMA>);
    if (e != cudaSuccess) return e;
    if (fa.numRegs * C::THREADS < C::NUM_EPI * 32 * C::EPI_REGS + (C::THREADS - C::NUM_EPI * 32) * C::AUX_REGS) return cudaErrorLaunchOutOfResources;
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3(num_sms / 2 * 2); q.blockDim = dim3(C::THREADS); q.dynamicSmemBytes = C::SMEM_BYTES;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    q.attrs = at; q.numAttrs = 1;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, k_edge_fused2_tc<M, TMA>, &q);
    if (e != cudaSuccess) return e;
    max_clusters = (n > 0 && n < num_sms / 2) ? n : num_sms / 2;
  }
  if (a.num_m_tiles % 2 != 0) return cudaErrorInvalidValue;
  const int units = a.num_m_tiles;                 // (num_m_tiles / 2) pairs x 2 slots
  if (units == 0) return cudaSuccess;
  const int nclusters = units < max_clusters ? units : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nclusters * 2); cfg.blockDim = dim3(C::THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = 2; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs; cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, k_edge_fused2_tc<M, TMA>, a, maps);
}

// Persistent variant: `nclusters` CTA pairs stay resident for all a.num_evals evaluations (no PDL attribute: nothing precedes it).
template <class M>
cudaError_t launch_fused2_persistent(const FusedArgs &a, int nclusters, cudaStream_t st) {
  using C = Fused2Cfg<M>;
  static bool configured_dev[64] = {};
  int dev_ = 0;
  cudaGetDevice(&dev_);
  if (!configured_dev[dev_ & 63]) {
    cudaError_t e = cudaFuncSetAttribute(k_edge_fused2_tc<M, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, k_edge_fused2_tc<M, false, true>);
    if (e != cudaSuccess) return e;
    if (fa.numRegs * C::THREADS < C::NUM_EPI * 32 * C::EPI_REGS + (C::THREADS - C::NUM_EPI * 32) * C::AUX_REGS) return cudaErrorLaunchOutOfResources;
    configured_dev[dev_ & 63] = true;
  }
  if (a.num_m_tiles % 2 != 0 || a.num_evals <= 0 || nclusters <= 0 || nclusters > a.num_m_tiles) return cudaErrorInvalidValue;
  static const PairMaps none = {};
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nclusters * 2); cfg.blockDim = dim3(C::THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = 2; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, k_edge_fused2_tc<M, false, true>, a, none);
}

// maps == nullptr selects the cp.async / relay variant
template <class M>
cudaError_t launch_fused2_tc(const FusedArgs &a, int num_sms, cudaStream_t st, const PairMaps *maps = nullptr) {
  if (maps) return launch_fused2_impl<M, true>(a, *maps, num_sms, st);
  static const PairMaps none = {};
  return launch_fused2_impl<M, false>(a, none, num_sms, st);
}

}  // namespace tc
}  // namespace ccsp
