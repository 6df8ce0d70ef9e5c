// kernels_fused3.cuh — EXPERIMENT (harness only, not used by the library): the CTA-pair edge kernel with the decoder's
// A operand in TENSOR MEMORY.
//
// Same computation and operand formats as k_edge_fused2_tc; what changes is how the first-layer activations reach GEMM2.
// fused2 writes them (hi / lo BF16) into a 64 KB shared-memory ring and GEMM2 reads them back three times as an SS MMA, and
// ncu shows that kernel limited by the shared-memory data pipe (profiles/README.md §3).  Here epilogue-1 stores the packed
// BF16 activations to tensor memory (tcgen05.st) and GEMM2 takes its A operand from there (TS-form cta_group::2 MMA):
//   * no shared-memory stores / proxy fences in epilogue-1, no A reads of GEMM2 from shared memory;
//   * the 64 KB of the operand ring become a fifth stage of ring 1 and a 6-deep decoder-weight ring;
//   * TMEM: D1 = columns 0..255, D2 = 256..383 (SINGLE-buffered), decoder operand slots = 384..511: 8 slots of 16 columns
//     (one k-step: 8 columns hi + 8 columns lo, two BF16 per column), two per epilogue column group;
//   * epilogue-2 runs in four warps of its own (a thread owns a whole row of D2: no cross-warp reduction).
//
// Result on B200 (csrc/tests/tc_gemm_test.cu pair): numerically identical to fused2; GEMM1 reaches the tensor-bound cadence
// (15.6 k cycles per unit incl. the interleaved GEMM2, vs 18.5 k in fused2) — but 0.127 ms vs 0.122 ms overall.  What blocks
// it is the TMEM budget, not the idea: D1 256 + operand slots 128 leave room for ONE D2, so GEMM2 of the next unit has to wait
// until epilogue-2 has drained D2 (8 k cycles with 56 registers per thread), GEMM2 completes 2-4 k cycles after its last
// k-step because it queues behind the next GEMM1 on the tensor pipe, and the stall propagates to epilogue-1 through the
// operand slots.  fused2's double-buffered D2 + deferred epilogue-2 hides exactly that latency.
#pragma once
#include "../kernels_fused2.cuh"

namespace ccsp {
namespace tc {

__device__ __forceinline__ void umma2_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
               "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}"
               ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 8 consecutive 32-bit columns, register -> TMEM
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <class M>
struct Fused3Cfg {
  static_assert(M::KIND == KIND_BF16, "BF16 operand modes only");
  static constexpr int NT1 = 256, NT2 = 128;
  static constexpr int B1_PART = (NT1 / 2) * ROWB;            // this CTA's 128 weight rows of one operand part: 8 KB
  static constexpr int B1_STAGE = M::NS * B1_PART;
  static constexpr int B1_BLOB_PART = NT1 * ROWB;
  static constexpr int B1_BLOB_STAGE = M::NS * B1_BLOB_PART;
  static constexpr int STAGE1 = M::A_STAGE + B1_STAGE;        // 32 KB (x3 split)
  static constexpr int NSTAGE1 = 5;
  static constexpr int W_PART = (NT2 / 2) * ROWB;             // this CTA's 64 decoder-weight rows of one part: 4 KB
  static constexpr int W_STAGE = M::NS * W_PART;
  static constexpr int W_BLOB_PART = NT2 * ROWB;
  static constexpr int W_BLOB_STAGE = M::NS * W_BLOB_PART;
  static constexpr int NW = 6;                                // of the 8 chunks a unit consumes: practically resident
  static constexpr int OFF_W = NSTAGE1 * STAGE1;
  static constexpr int OFF_EXTRA = OFF_W + NW * W_STAGE;
  static constexpr int NUM_EPI = 16, EPI_T = NUM_EPI * 32;    // epilogue-1 warps
  static constexpr int WARP_EPI2_0 = NUM_EPI, NUM_EPI2 = 4;   // epilogue-2 warps (one per TMEM lane quarter)
  static constexpr int WARP_PROD0 = WARP_EPI2_0 + NUM_EPI2, NUM_PROD_WARPS = 4;
  static constexpr int WARP_LOAD = 24, WARP_MMA1 = 25, WARP_MMA2 = 26, WARP_LOADW = 27;
  static constexpr int THREADS = 28 * 32;                     // 896
  // register pool = 896 x 72: 512 x 96 (epilogue-1) + 128 x 56 (epilogue-2) + 128 x 32 (gather) + 128 x 32 (loaders, issuers)
  static constexpr int EPI_REGS = 96, EPI2_REGS = 56, AUX_REGS = 32;
  // barriers 512 | tb_s 2x256 | bd1 128 | w2t 128x8 | bd2 16
  static constexpr int SMEM_EXTRA = 512 + (512 + CCSP_HH + CCSP_MAXP * CCSP_HH + 16) * 4;
  static constexpr int SMEM_BYTES = OFF_EXTRA + SMEM_EXTRA + 1024;
  static constexpr int D2_COL = 256, A2_COL = 384;
  static_assert(SMEM_BYTES <= 227 * 1024, "kernel does not fit in shared memory");
};

template <class M>
__global__ void __launch_bounds__(Fused3Cfg<M>::THREADS, 1) k_edge_fused3_tc(const FusedArgs A) {
  using C = Fused3Cfg<M>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t *extra = smem + C::OFF_EXTRA;
  uint64_t *full1 = reinterpret_cast<uint64_t *>(extra);     // [8] stage of ring 1 filled (leader: incl. the peer's)
  uint64_t *empty1 = full1 + 8;                              // [8]
  uint64_t *a2_full = empty1 + 8;                            // [4 groups][2 slots] decoder operand k-step written (leader only)
  uint64_t *a2_empty = a2_full + 8;                          // [4][2]
  uint64_t *w_full = a2_empty + 8;                           // [8]
  uint64_t *w_empty = w_full + 8;                            // [8]
  uint64_t *tfull1 = w_empty + 8, *tempty1 = tfull1 + 1;     // D1
  uint64_t *tfull2 = tempty1 + 1, *tempty2 = tfull2 + 1;     // D2
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tempty2 + 1);
  float *tb_s = reinterpret_cast<float *>(extra + 512);      // [2][256]
  float *bd1 = tb_s + 512;                                   // [128]
  float *w2t = bd1 + CCSP_HH;                                // [128][8]
  float *bd2 = w2t + CCSP_MAXP * CCSP_HH;                    // [8] (+8 pad)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int num_units = (A.num_m_tiles / 2) * 2;             // (pair of 128-edge tiles, slot)
  const int unit0 = blockIdx.x >> 1, unit_step = gridDim.x >> 1;
  const uint32_t smem_base = smem_u32(smem);
  long long *const tr = (A.trace && blockIdx.x == 0) ? A.trace : nullptr;
#define TR(role, slot) do { if (tr && it < 8) tr[((role) * 8 + it) * 16 + (slot)] = clock64(); } while (0)

  if (threadIdx.x == 0) pdl_launch_dependents();
  if (threadIdx.x == 0) {
    // ring 1: own gather threads + own weight loader (+ at the leader: the peer's relay lane)
    for (int s = 0; s < C::NSTAGE1; ++s) { mbar_init(&full1[s], C::NUM_PROD_WARPS * 32 + 1 + (leader ? 1 : 0)); mbar_init(&empty1[s], 1); }
    for (int s = 0; s < 8; ++s) { mbar_init(&a2_full[s], 8); mbar_init(&a2_empty[s], 1); }
    for (int s = 0; s < C::NW; ++s) { mbar_init(&w_full[s], leader ? 2 : 1); mbar_init(&w_empty[s], 1); }
    mbar_init(tfull1, 1); mbar_init(tempty1, 2 * C::NUM_EPI);
    mbar_init(tfull2, 1); mbar_init(tempty2, 2 * C::NUM_EPI2);
    fence_barrier_init();
  }
  if (warp == C::WARP_MMA1) tmem_alloc2(tmem_ptr, 512);
  if (warp < C::NUM_EPI) {     // tables of epilogue-2
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

#define REG_DEC() asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::AUX_REGS))
#define REG_INC() asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C::EPI_REGS))
#define REG_DEC2() asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(C::EPI2_REGS))

  if (warp >= C::WARP_EPI2_0 && warp < C::WARP_EPI2_0 + C::NUM_EPI2) {
    REG_DEC2();
    // ============ epilogue-2 (own warps): D2 -> + b -> SiLU -> 128 -> P -> o.  A thread owns a whole row, so there is no
    // cross-warp reduction; D2 is single-buffered, and these warps drain it while GEMM1 of the next unit runs, long before
    // GEMM2 of that unit can start.
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t r_tempty2 = mapa_u32(smem_u32(tempty2), 0);
    uint32_t it = 0;
    for (int u = unit0; u < num_units; u += unit_step, ++it) {
      const int mt = (u >> 1) * 2 + (int)rank, slot = u & 1;
      const size_t row = (size_t)mt * SUB_M + r;
      mbar_wait_cl(tfull2, it & 1);
      if (it == 0) pdl_wait();               // o is still being read by the preceding node kernel until it completes
      tc_fence_after();
      const uint32_t taddr2 = tmem_base + C::D2_COL + ((uint32_t)(quarter * 32) << 16);
      float acc[CCSP_MAXP];
#pragma unroll
      for (int p = 0; p < CCSP_MAXP; ++p) acc[p] = bd2[p];
#pragma unroll 1
      for (int hc = 0; hc < 8; ++hc) {
        float v2[16];
        tmem_ld16(taddr2 + hc * 16, v2);
        if (hc == 7) {                       // D2 fully read: GEMM2 of the next unit may overwrite it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(r_tempty2);
        }
        const int c0 = hc * 16;
        if (A.P <= 4) {
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const float d = silu_raw(v2[jj] + bd1[c0 + jj]);
            const float4 w = *reinterpret_cast<const float4 *>(&w2t[(c0 + jj) * CCSP_MAXP]);
            acc[0] = fmaf(d, w.x, acc[0]); acc[1] = fmaf(d, w.y, acc[1]); acc[2] = fmaf(d, w.z, acc[2]); acc[3] = fmaf(d, w.w, acc[3]);
          }
        } else {
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            const float d = silu_raw(v2[jj] + bd1[c0 + jj]);
            const float4 w0 = *reinterpret_cast<const float4 *>(&w2t[(c0 + jj) * CCSP_MAXP]);
            const float4 w1 = *reinterpret_cast<const float4 *>(&w2t[(c0 + jj) * CCSP_MAXP + 4]);
            acc[0] = fmaf(d, w0.x, acc[0]); acc[1] = fmaf(d, w0.y, acc[1]); acc[2] = fmaf(d, w0.z, acc[2]); acc[3] = fmaf(d, w0.w, acc[3]);
            acc[4] = fmaf(d, w1.x, acc[4]); acc[5] = fmaf(d, w1.y, acc[5]); acc[6] = fmaf(d, w1.z, acc[6]); acc[7] = fmaf(d, w1.w, acc[7]);
          }
        }
      }
      if (!(A.dbg & 4)) {
        float *orow = A.o + ((size_t)row * 2 + slot) * A.P;
        if (A.P == 4) {
          *reinterpret_cast<float4 *>(orow) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        } else {
#pragma unroll
          for (int p = 0; p < CCSP_MAXP; ++p)
            if (p < A.P) orow[p] = acc[p];
        }
      }
    }
  } else if (warp >= C::WARP_PROD0 && warp < C::WARP_PROD0 + C::NUM_PROD_WARPS) {
    REG_DEC();
    // ============ A gather (as in k_edge_fused2_tc): this CTA's 128 edges =================================
    const int t = threadIdx.x - C::WARP_PROD0 * 32;
    constexpr int LPR = M::NS == 2 ? 8 : 4;        // lanes per row
    constexpr int RPP = 128 / LPR;                 // rows per pass
    constexpr int NP = SUB_M / RPP;                // passes (= cp.async per thread and chunk)
    const int q8 = t % LPR, r0 = t / LPR;
    const int part = q8 >> 2, q = q8 & 3;
    uint32_t g = 0;
    pdl_wait();                          // pe_split is written by the preceding node kernel
    for (int u = unit0; u < num_units; u += unit_step) {
      const int m0 = ((u >> 1) * 2 + (int)rank) * SUB_M;
      uint32_t roff[NP];
#pragma unroll 1
      for (int kc = 0; kc < M::NKC1; ++kc, ++g) {
        if (kc == 0 || kc == M::NKC1 / 2) {
          const int *idx = kc == 0 ? A.src_i : A.src_j;
#pragma unroll
          for (int p = 0; p < NP; ++p) roff[p] = (uint32_t)__ldg(&idx[m0 + r0 + RPP * p]) * (M::PE_ROW_BYTES / 16);
        }
        const uint32_t s = g % C::NSTAGE1;
        mbar_wait(&empty1[s], ((g / C::NSTAGE1) & 1) ^ 1);
        if (!(A.dbg & 1)) {
          const uint32_t koff = M::pe_off(kc % (M::NKC1 / 2), part, q);
          const uint32_t st = smem_base + s * C::STAGE1 + part * PART;
#pragma unroll
          for (int p = 0; p < NP; ++p)
            cp_async16(st + sw64_off(r0 + RPP * p, q), A.pe_split + (size_t)roff[p] * 16 + koff);
        }
        cp_async_arrive_noinc(&full1[s]);
      }
    }
  } else if (warp >= C::WARP_LOAD) {
    REG_DEC();       // one instruction for the whole warpgroup (warps 20-23), then the per-warp roles
    if (warp == C::WARP_LOAD) {
      if (lane == 0) {
        // ============ first-layer weights: this CTA's 128 of the 256 rows of every chunk =================
        uint32_t g = 0;
        for (int u = unit0; u < num_units; u += unit_step) {
          const int mt = (u >> 1) * 2, slot = u & 1;
          const int grp = __ldg(&A.tile_type[mt]);
          const uint8_t *blob = A.b_blob + ((size_t)(grp * 2 + slot) * M::NKC1) * C::B1_BLOB_STAGE + rank * C::B1_PART;
          for (int kc = 0; kc < M::NKC1; ++kc, ++g) {
            const uint32_t s = g % C::NSTAGE1;
            mbar_wait(&empty1[s], ((g / C::NSTAGE1) & 1) ^ 1);
            if (A.dbg & 2) { mbar_arrive(&full1[s]); continue; }
            mbar_arrive_expect_tx(&full1[s], C::B1_STAGE);
            const uint32_t dst = smem_base + s * C::STAGE1 + M::A_STAGE;
            const uint8_t *src = blob + (size_t)kc * C::B1_BLOB_STAGE;
            bulk_g2s(dst, src, C::B1_PART, &full1[s]);
            if (M::NS == 2) bulk_g2s(dst + C::B1_PART, src + C::B1_BLOB_PART, C::B1_PART, &full1[s]);
          }
        }
      }
    } else if (warp == C::WARP_LOADW) {
      if (lane == 0) {
        // ============ decoder weights: this CTA's 64 of the 128 rows, chunks in GEMM2's consumption order ==
        uint32_t g2 = 0;
        const uint8_t *blob = A.w_blob + rank * C::W_PART;
        for (int u = unit0; u < num_units; u += unit_step) {
          for (int q = 0; q < 8; ++q, ++g2) {
            const uint32_t s = g2 % C::NW;
            mbar_wait(&w_empty[s], ((g2 / C::NW) & 1) ^ 1);
            if (A.dbg & 2) { mbar_arrive(&w_full[s]); continue; }
            const int c = 2 * (q & 3) + (q >> 2);
            mbar_arrive_expect_tx(&w_full[s], C::W_STAGE);
            const uint32_t dst = smem_base + C::OFF_W + s * C::W_STAGE;
            const uint8_t *src = blob + (size_t)c * C::W_BLOB_STAGE;
            bulk_g2s(dst, src, C::W_PART, &w_full[s]);
            if (M::NS == 2) bulk_g2s(dst + C::W_PART, src + C::W_BLOB_PART, C::W_PART, &w_full[s]);
          }
        }
      }
    } else if (warp == C::WARP_MMA1) {
      if (lane == 0) {
        if (leader) {
          // ============ GEMM1 issuer (M = 256 over the pair) =============================================
          uint32_t g = 0, it = 0;
          for (int u = unit0; u < num_units; u += unit_step, ++it) {
            TR(0, 0);
            mbar_wait_cl(tempty1, (it & 1) ^ 1);
            TR(0, 1);
            tc_fence_after();
            bool ready = false;
            for (int kc = 0; kc < M::NKC1; ++kc, ++g) {
              const uint32_t s = g % C::NSTAGE1;
              if (!ready) mbar_wait_cl(&full1[s], (g / C::NSTAGE1) & 1);
              if (kc == 0) TR(0, 2);
              if (kc == 8) TR(0, 3);
              tc_fence_after();
              const uint32_t a_hi = smem_base + s * C::STAGE1;
              if (!(A.dbg & 8)) issue_kstep2<M, C::NT1>(tmem_base, a_hi, a_hi + M::A_STAGE, C::B1_PART, 0, kc == 0);
              ready = kc + 1 < M::NKC1 && mbar_test_wait(&full1[(g + 1) % C::NSTAGE1], ((g + 1) / C::NSTAGE1) & 1);
              if (!(A.dbg & 8)) issue_kstep2<M, C::NT1>(tmem_base, a_hi, a_hi + M::A_STAGE, C::B1_PART, 1, false);
              umma_commit2(&empty1[s]);
            }
            umma_commit2(tfull1);
            TR(0, 4);
          }
        } else {
          // ============ peer: forward "stage s is full here (A rows + weight half)" to the leader ==========
          uint32_t g = 0;
          for (int u = unit0; u < num_units; u += unit_step) {
            for (int kc = 0; kc < M::NKC1; ++kc, ++g) {
              const uint32_t s = g % C::NSTAGE1;
              mbar_wait(&full1[s], (g / C::NSTAGE1) & 1);
              mbar_arrive_remote(mapa_u32(smem_u32(&full1[s]), 0));
            }
          }
        }
      }
    } else if (warp == C::WARP_MMA2) {
      if (lane == 0) {
        if (leader) {
          // ============ GEMM2 issuer: A from tensor memory ================================================
          constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C::NT2 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
          uint32_t g2 = 0, it = 0;
          const uint32_t d2 = tmem_base + C::D2_COL;
          for (int u = unit0; u < num_units; u += unit_step, ++it) {
            TR(1, 0);
            mbar_wait_cl(tempty2, (it & 1) ^ 1);      // epilogue-2 of the previous unit has drained D2
            TR(1, 1);
            tc_fence_after();
            for (int q = 0; q < 8; ++q, ++g2) {
              const uint32_t cg = q & 3, j = q >> 2, ws = g2 % C::NW;
              mbar_wait_cl(&w_full[ws], (g2 / C::NW) & 1);
              const uint32_t b_hi = smem_base + C::OFF_W + ws * C::W_STAGE, b_lo = b_hi + C::W_PART;
#pragma unroll
              for (int par = 0; par < 2; ++par) {    // k-step 2 j + par of the group's four: slot (cg, par)
                mbar_wait_cl(&a2_full[2 * cg + par], (it * 2 + j) & 1);
                if (q == 0 && par == 0) TR(1, 2);
                tc_fence_after();
                const uint32_t a_hi = tmem_base + C::A2_COL + (2 * cg + par) * 16, a_lo = a_hi + 8;
                const uint32_t ko = par * 32;
                if (!(A.dbg & 8)) {
                  const uint32_t acc = !(q == 0 && par == 0);
                  if (M::NSPLIT == 3) {
                    umma2_f16_ts(d2, a_lo, smem_desc_sw64(b_hi + ko), idesc, acc);
                    umma2_f16_ts(d2, a_hi, smem_desc_sw64(b_lo + ko), idesc, 1);
                    umma2_f16_ts(d2, a_hi, smem_desc_sw64(b_hi + ko), idesc, 1);
                  } else {
                    umma2_f16_ts(d2, a_hi, smem_desc_sw64(b_hi + ko), idesc, acc);
                  }
                }
                umma_commit2(&a2_empty[2 * cg + par]);
              }
              umma_commit2(&w_empty[ws]);
            }
            umma_commit2(tfull2);
            TR(1, 6);
          }
        } else {
          uint32_t g2 = 0;
          for (int u = unit0; u < num_units; u += unit_step) {
            for (int q = 0; q < 8; ++q, ++g2) {
              const uint32_t s = g2 % C::NW;
              mbar_wait(&w_full[s], (g2 / C::NW) & 1);
              mbar_arrive_remote(mapa_u32(smem_u32(&w_full[s]), 0));
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
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    const uint32_t r_tempty1 = mapa_u32(smem_u32(tempty1), 0);
    const uint32_t r_a2_full0 = mapa_u32(smem_u32(&a2_full[2 * cg]), 0), r_a2_full1 = mapa_u32(smem_u32(&a2_full[2 * cg + 1]), 0);
    const uint64_t pol_s = l2_policy_evict_first();
    uint32_t it = 0;
    for (int u = unit0; u < num_units; u += unit_step, ++it) {
      const int mt = (u >> 1) * 2 + (int)rank, slot = u & 1;
      const int grp = __ldg(&A.tile_type[mt]);
      const size_t row = (size_t)mt * SUB_M + r;
      const int gcol0 = slot * 256 + cg * 64;
      // S in the blocked layout: 32-row x 32-col blocks of 8 pieces x 32 lanes x 16 B (coalesced 512 B / instruction)
      const float4 *Sblk = reinterpret_cast<const float4 *>(A.S) + (((row >> 5) * 16 + (gcol0 >> 5)) * 8) * 32 + lane;
      const bool noS = (A.dbg & 4) != 0;
      float4 sa = noS ? make_float4(0.f, 0.f, 0.f, 0.f) : ldg_nc_f4_hint(Sblk, pol_s), sb = noS ? sa : ldg_nc_f4_hint(Sblk + 32, pol_s);
      if (threadIdx.x < 4 && u + unit_step < num_units && !(A.dbg & 4)) {
        const int un = u + unit_step;      // next unit's slice of S (4 row blocks x 32 KB contiguous) -> L2
        const size_t rb = (size_t)((un >> 1) * 2 + (int)rank) * 4 + threadIdx.x;
        prefetch_l2_bulk(A.S + (rb * 16 + (size_t)(un & 1) * 8) * 1024, 32768, pol_s);
      }
      float *tbu = tb_s + (it & 1) * 256;
      if (threadIdx.x < 256) tbu[threadIdx.x] = __ldg(&A.tb[(size_t)grp * CCSP_H2 + slot * 256 + threadIdx.x]);
      const int trole = warp == 0 ? 2 : 3;
      const bool tron = lane == 0 && (warp == 0 || warp == 12);
      if (tron) TR(trole, 0);
      asm volatile("bar.sync 1, 512;" ::: "memory");
      // ---- epilogue-1: D1 -> registers -> SiLU -> BF16 hi / lo -> tensor memory (decoder operand k-steps) ---------
      mbar_wait_cl(tfull1, it & 1);
      if (tron) TR(trole, 2);
      tc_fence_after();
      const uint32_t taddr1 = tmem_base + cg * 64 + lane_sel;
      float vall[64];
      tmem_ld32(taddr1, vall);
      tmem_ld32(taddr1 + 32, vall + 32);
      tc_fence_before();                             // D1 fully in registers: GEMM1 of the next unit may overwrite it
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(r_tempty1);
      if (tron) TR(trole, 3);
#pragma unroll
      for (int ks4 = 0; ks4 < 4; ++ks4) {            // 16 columns = one k-step of the decoder
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {
          const int pi = ks4 * 2 + h8;
          const float4 ca = sa, cb = sb;
          if (pi < 7 && !noS) {                      // S of the next piece, one step ahead
            const float4 *nx = Sblk + ((pi + 1) >> 2) * 256 + ((pi + 1) & 3) * 64;
            sa = ldg_nc_f4_hint(nx, pol_s); sb = ldg_nc_f4_hint(nx + 32, pol_s);
          }
          const float4 ta = *reinterpret_cast<const float4 *>(&tbu[cg * 64 + pi * 8]);
          const float4 tb4 = *reinterpret_cast<const float4 *>(&tbu[cg * 64 + pi * 8 + 4]);
          const float *v = vall + pi * 8;
          float f[8];
          f[0] = silu_raw(v[0] + ca.x + ta.x); f[1] = silu_raw(v[1] + ca.y + ta.y);
          f[2] = silu_raw(v[2] + ca.z + ta.z); f[3] = silu_raw(v[3] + ca.w + ta.w);
          f[4] = silu_raw(v[4] + cb.x + tb4.x); f[5] = silu_raw(v[5] + cb.y + tb4.y);
          f[6] = silu_raw(v[6] + cb.z + tb4.z); f[7] = silu_raw(v[7] + cb.w + tb4.w);
          split_pair(f[0], f[1], hi[h8 * 4 + 0], lo[h8 * 4 + 0]); split_pair(f[2], f[3], hi[h8 * 4 + 1], lo[h8 * 4 + 1]);
          split_pair(f[4], f[5], hi[h8 * 4 + 2], lo[h8 * 4 + 2]); split_pair(f[6], f[7], hi[h8 * 4 + 3], lo[h8 * 4 + 3]);
        }
        const int par = ks4 & 1;
        // GEMM2 is done with this slot (its previous k-step): use count of slot (cg, par) = 2 per unit
        mbar_wait(&a2_empty[2 * cg + par], ((it * 2 + (ks4 >> 1)) & 1) ^ 1);
        tc_fence_after();
        const uint32_t ta2 = tmem_base + C::A2_COL + (2 * cg + par) * 16 + lane_sel;
        tmem_st8(ta2, hi);
        if (M::NS == 2) tmem_st8(ta2 + 8, lo);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(par ? r_a2_full1 : r_a2_full0);
        if (tron) TR(trole, 4 + ks4);
      }
      if (tron) TR(trole, 10);
    }
  }
#undef TR
#undef REG_DEC
#undef REG_INC
  tc_fence_before();
  __syncthreads();
  cluster_sync();                      // no CTA exits (or frees TMEM) while its peer may still signal it or read its operands
  if (warp == C::WARP_MMA1) tmem_dealloc2(tmem_base, 512);
}

template <class M>
cudaError_t launch_fused3_tc(const FusedArgs &a, int num_sms, cudaStream_t st) {
  using C = Fused3Cfg<M>;
  static int max_clusters_dev[64] = {};      // per device: function attributes and cluster occupancy
  int dev_ = 0;
  cudaGetDevice(&dev_);
  int &max_clusters = max_clusters_dev[dev_ & 63];
  if (max_clusters == 0) {
    cudaError_t e = cudaFuncSetAttribute(k_edge_fused3_tc<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    // setmaxnreg moves registers inside the pool the CTA was launched with (see launch_fused2_impl)
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, k_edge_fused3_tc<M>);
    if (e != cudaSuccess) return e;
    if (fa.numRegs * C::THREADS < C::NUM_EPI * 32 * C::EPI_REGS + C::NUM_EPI2 * 32 * C::EPI2_REGS + 8 * 32 * C::AUX_REGS) return cudaErrorLaunchOutOfResources;
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3(num_sms / 2 * 2); q.blockDim = dim3(C::THREADS); q.dynamicSmemBytes = C::SMEM_BYTES;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    q.attrs = at; q.numAttrs = 1;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, k_edge_fused3_tc<M>, &q);
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
  return cudaLaunchKernelEx(&cfg, k_edge_fused3_tc<M>, a);
}

}  // namespace tc
}  // namespace ccsp
