// Skeleton shared by the fused tcgen05 MLP kernels (forward: mlp_tc.cu, dgrad: mlp_tc_bwd.cu).
//
// A "chain" is a sequence of GEMM steps  D_s = X_s . W_s^T  over ONE 128-row sample tile whose intermediate
// activations never leave the SM:
//   * accumulators are DOUBLE-BUFFERED in tensor memory (2 x 256 fp32 columns, alternating by step): the tensor
//     core works on step s+1 while the epilogue warps are still draining step s;
//   * the fp16 activations that feed the next step live in ONE tile [128 x 256] in TENSOR memory (written in place
//     into the accumulator just drained: every epilogue warp packs its 32-column chunk into the first 16 columns of
//     that same chunk, so nobody overwrites columns another warp has yet to load; read by TS-mode MMAs; no hazard: the
//     epilogue of step s starts only after every MMA of step s has completed).  ChRing::ACT_IN_TMEM = false selects the older
//     shared-memory activation tile in the UMMA operand layout;
//   * the rest of shared memory is a ring (ChRing: 4 x 40 KB forward, 4 x 32 KB dgrad) of weight blocks [N x 64 K] streamed with bulk async
//     copies (TMA engine) in exactly the order the MMAs consume them; the bias rides in the MMA as one extra
//     K=16 slice against a constant ones column.
// Hand-off between steps is per 64 K-values: the eight epilogue warps drain the accumulator 64 columns at a time
// (tcgen05.ld -> activation -> fp16 -> st.shared; both warpgroups share each group, 32 columns each, because the
// fp32->fp16 pack instruction is the slow part) and publish each 64-column group of the next step's A operand on
// its own mbarrier, so step s+1 starts on group 0 while groups 1..3 are still being converted.
//
// Issue budget (tools/bench_umma.py): one thread needs ~43 cycles per tcgen05.mma plus ~130 per mbarrier wait;
// an N=256 K=16 MMA occupies the tensor pipe for 128 cycles.  Blocks of 4-5 MMAs are dealt alternately to TWO
// issuer warps, which keeps each of them well below the pipe's pace.
//
// Warp roles (512 threads, 1 CTA per SM, persistent over tiles):
//   warps 0-3     prologue warpgroup: builds the NEXT tile's shared-memory operands (positional encodings /
//                 head-gradient tile) while the current tile is in flight
//   warps 4-7     epilogue warpgroup 0  (columns [0,32) of every 64-column output group)
//   warps 8-11    epilogue warpgroup 1  (columns [32,64) of every group)
//   warp 12       weight producer (one lane)
//   warps 14, 15  MMA issuers (converged warps, one elected lane issues): warp 14+i issues blocks i, i+2, ... of
//                 every step; tokens (ChainBars::issued) make the blocks of a step enter the tensor pipe in order.
//                 Warp 14 also owns the TMEM allocation.
// The SM's warp arbiter favours the highest warp id among eligible warps: the latency-critical single-thread
// issuers sit at the top, the ALU-heavy but latency-tolerant encoder warps at the bottom.
#pragma once
#include <stdlib.h>
#include "mlp_tc.cuh"

namespace lush {

constexpr int CH_THREADS = 512;
constexpr int CH_MAX_NSTAGE = 4;
// Weight ring, chosen per kernel (template parameter R of the producer / issuer):
//   forward: 4 stages x 40 KB -- a [256 x 64] weight block + the [256 x 16] bias block ride in one stage (the
//            forward kernel keeps its activations in tensor memory, so the 64 KB activation tile is not allocated)
//   dgrad:   4 stages x 32 KB -- no bias slices on the trunk steps and ~30 KB of spare shared memory: with as many
//            stages as blocks per step, block j of a step reuses the stage of block j of the PREVIOUS step, so every
//            copy is requested a whole step ahead (the copies take ~3 000 cycles under load; with 3 stages the last
//            block of every step arrived ~2 000 cycles after its activations were ready)
template <int NSTAGE_, int STAGE_BYTES_, bool ACT_IN_TMEM_ = false>
struct ChRing {
  // true: the epilogue writes the next step's A operand into TENSOR memory (in place, inside the
  // accumulator it has just drained) and the hidden-block MMAs read it from there (TS mode): no activation stores /
  // operand reads / proxy fences on the shared-memory port, which is the contended resource of these kernels
  static constexpr bool ACT_IN_TMEM = ACT_IN_TMEM_;
  static constexpr int NSTAGE = NSTAGE_;
  static constexpr int STAGE_BYTES = STAGE_BYTES_;
  static constexpr int STAGE0 = 0;
  static constexpr int ACT = STAGE0 + NSTAGE_ * STAGE_BYTES_;        // [128 x 256] activations (A operand)
  static constexpr int FREE0 = ACT + (ACT_IN_TMEM_ ? 0 : 128 * 256 * 2);   // kernel-specific operands from here (ChainMap)
  static_assert(NSTAGE_ >= 2 && NSTAGE_ <= CH_MAX_NSTAGE, "ring depth");
};
using FwdRing = ChRing<4, 256 * (64 + 16) * 2, true>;   // the 64 KB activation tile is gone: one more 40 KB stage
using DgradRing = ChRing<4, 256 * 64 * 2, true>;
constexpr int CH_MAX_STEPS = kMaxDepth + 6;
constexpr int CH_WG_PRO = 0;                   // first warp of the prologue warpgroup
constexpr int CH_WG_EPI0 = 4;                  // first warp of epilogue warpgroup 0
constexpr int CH_WARP_PRODUCER = 12;
constexpr int CH_WARP_MMA0 = 14;               // issuer of even blocks (odd blocks: CH_WARP_MMA0 + 1)
constexpr int CH_X16_BYTES = TROWS * 16 * 2;   // [128 x 16] fp16 tile

// byte offsets of the kernel-specific shared-memory operands (strides: per tile parity b; 0 = single buffer)
struct ChainMap {
  uint32_t enc, encv, encv_stride, x16, x16_stride;
};

// One GEMM step.  Its K range is, in this order: an optional K=16 slice against the X16 shared-memory tile
// (bias ones / head gradients), optional K=32 slices against an encoding tile, and h_slices K=32 slices against
// the activation tile (h_slices is even).
// Weight stream per step: [pre16 block N x 16][enc block N x 32*enc_slices][hidden blocks N x 64 ...], each in
// the tile-canonical layout with N rows.  Ring block 0 of a step = pre16 + enc (+ the first hidden block when
// enc_slices < 2 and the sum fits one ring stage); every further hidden block is one ring block.
struct ChainStep {
  int16_t N;          // output columns (256 | 128 | 64)
  int8_t pre16;       // 0 | 1
  int8_t enc_slices;  // 0 | 1 | 2
  int8_t enc_src;     // 0: ENC (64 columns), 1: ENCV (32 columns)
  int8_t h_slices;    // 0 | 4 | 8
  int8_t wait_a;      // 1: the activation tile is being written by the previous step's epilogue (wait per group)
  int8_t writes_a;    // 1: the epilogue writes the activation tile (and publishes its 64-column groups)
  int8_t free_enc;    // after this step's block 0: 1 = ENC, 2 = ENCV / X16 of this tile are no longer read
  int8_t epi;         // kernel-specific epilogue kind
  int8_t keep_acc;    // 1: the NEXT step uses the same accumulator buffer again (its A operand, produced two steps ago,
                      //    lives in the other one); 0: the buffers alternate
  int16_t p0, p1;     // epilogue parameters
  int32_t off0, off1; // epilogue byte offsets
};

struct ChainPlan {
  int nsteps;
  int debug;              // LUSH_CHAIN_DEBUG: bit 0 = do not stream weights (timing experiments only, results are
                          // garbage), bit 1 = record the timeline trace of CTA 0 (tools/trace_chain.py)
  int64_t stream_bytes;   // weight stream per tile
  ChainStep steps[CH_MAX_STEPS];
};

__host__ __device__ inline int chain_merge_h0(const ChainStep& s, int stage_bytes) {
  return (s.enc_slices < 2 && s.h_slices > 0 && s.N * 2 * (16 * s.pre16 + KS * s.enc_slices + 64) <= stage_bytes) ? 1 : 0;
}
__host__ __device__ inline int chain_block0_bytes(const ChainStep& s, int stage_bytes) {
  return s.N * 2 * (16 * s.pre16 + KS * s.enc_slices + 64 * chain_merge_h0(s, stage_bytes));
}
__host__ __device__ inline int chain_nblocks(const ChainStep& s, int stage_bytes) {
  return 1 + (s.h_slices >> 1) - chain_merge_h0(s, stage_bytes);
}
__host__ __device__ inline int chain_step_bytes(const ChainStep& s) {
  return s.N * 2 * (16 * s.pre16 + KS * (s.enc_slices + s.h_slices));
}
// returns false if a ring block of the plan does not fit one stage of the kernel's ring
static inline bool chain_finish_plan(ChainPlan& P, int stage_bytes) {
  const char* dbg = getenv("LUSH_CHAIN_DEBUG");
  P.debug = dbg ? atoi(dbg) : 0;
  int64_t b = 0;
  bool fits = true;
  for (int s = 0; s < P.nsteps; ++s) {
    b += chain_step_bytes(P.steps[s]);
    fits = fits && chain_block0_bytes(P.steps[s], stage_bytes) <= stage_bytes && P.steps[s].N * 64 * 2 <= stage_bytes;
  }
  P.stream_bytes = b;
  return fits;
}

// ---- optional timeline trace (LUSH_CHAIN_DEBUG bit 1): CTA 0 records clock64() of its third tile's events
//   slot = step*16 + {0,1: issuer 0 start/end; 2,3: issuer 1; 4 + 4*g + {0: d_ready seen, 1: accumulator released,
//   2,3: round 2g / round 2g+1 published} for epilogue warpgroup g; 12..15: the step's last weight block: issuer
//   reaches it, weights landed, activation group ready, MMAs issued}
// The probes sit on the epilogue's critical path (~4 % of the forward kernel), so they are compiled in only with
// -DLUSH_CHAIN_TRACE (LUSH_CHAIN_TRACE=1 python -m lush_nerf_b200.build --force; tools/trace_chain.py).
static __device__ long long g_chain_trace[CH_MAX_STEPS * 16];   // one copy per translation unit
__device__ __forceinline__ void chain_trace(const ChainPlan& P, int64_t it, int step, int ev) {
#ifdef LUSH_CHAIN_TRACE
  if ((P.debug & 2) && blockIdx.x == 0 && it == 2 && (threadIdx.x & 31) == 0) g_chain_trace[step * 16 + ev] = clock64();
#endif
}

struct ChainBars {
  uint64_t full[CH_MAX_NSTAGE], empty[CH_MAX_NSTAGE];
  uint64_t a_ready[4];          // K-values [64j, 64j+64) of the activation tile are complete (8 warps)
  uint64_t d_ready[2];          // accumulator d complete (one commit per issuer)
  uint64_t d_free[2];           // accumulator d is in the epilogue warps' registers (8 warps)
  uint64_t issued[2];           // issuer i has issued one more of its blocks: block k of a step enters the tensor pipe
                                // after block k-1 (the other issuer's), so every accumulator sums its K slices in ONE order
  uint64_t pro_ready[2];        // prologue operands of tile parity b are in shared memory (4 warps)
  uint64_t enc_free;            // the single ENC tile is no longer read by this tile's MMAs (commit)
  uint64_t pro_free[2];         // ENCV / X16 tiles of parity b are no longer read (commit)
  uint64_t aux;                 // kernel-specific (forward: head partial sums of warpgroup 0, 4 warps)
};

__device__ __forceinline__ void chain_init_bars(ChainBars& B) {
  for (int s = 0; s < CH_MAX_NSTAGE; ++s) { mbar_init(&B.full[s], 1); mbar_init(&B.empty[s], 1); }
  for (int j = 0; j < 4; ++j) mbar_init(&B.a_ready[j], 8);
  for (int r = 0; r < 2; ++r) {
    mbar_init(&B.d_ready[r], 2); mbar_init(&B.d_free[r], 8);
    mbar_init(&B.pro_ready[r], 4); mbar_init(&B.pro_free[r], 1);
  }
  mbar_init(&B.enc_free, 1);
  mbar_init(&B.issued[0], 1); mbar_init(&B.issued[1], 1);
  mbar_init(&B.aux, 4);
  fence_barrier_init();
}

// ---- warp 12, one lane: stream the packed weight blocks of every tile through the ring
template <class R>
__device__ __forceinline__ void chain_producer(ChainBars& B, uint8_t* smem, const uint8_t* __restrict__ wstream,
                                               const ChainPlan& P, int64_t ntiles) {
  int stage = 0; uint32_t phase = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const uint8_t* src = wstream;
    for (int s = 0; s < P.nsteps; ++s) {
      const ChainStep& S = P.steps[s];
      const uint32_t first = (uint32_t)chain_block0_bytes(S, R::STAGE_BYTES), big = (uint32_t)S.N * 64 * 2;
      const int nblk = chain_nblocks(S, R::STAGE_BYTES);
      for (int j = 0; j < nblk; ++j) {
        const uint32_t bytes = j == 0 ? first : big;
        mbar_wait(&B.empty[stage], phase ^ 1);
        if (P.debug & 1) { mbar_arrive(&B.full[stage]); }
        else {
          mbar_arrive_expect_tx(&B.full[stage], bytes);
          bulk_g2s(smem + R::STAGE0 + stage * R::STAGE_BYTES, src, bytes, &B.full[stage]);
        }
        src += bytes;
        if (++stage == R::NSTAGE) { stage = 0; phase ^= 1; }
      }
    }
  }
}

// ---- MMA issue blocks: NK16 consecutive K=16 MMAs with both operands in shared memory (descriptor low words
// a_lo / b_lo advance by ak16 / bk16 per K=16; both share the high word `hi`).  Executed by one elected lane.
#define LUSH_SS_FIRST                                                       \
  "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"       \
  "setp.ne.b32 p, %7, 0;\n\t"                                               \
  "mov.b32 al, %1;\n\tmov.b32 bl, %2;\n\t"                                  \
  "mov.b64 da, {al, %3};\n\tmov.b64 db, {bl, %3};\n\t"                      \
  "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, p;\n\t"            \
  "setp.eq.b32 p, %7, %7;\n\t"
#define LUSH_SS_NEXT                                                        \
  "add.u32 al, al, %4;\n\tadd.u32 bl, bl, %5;\n\t"                          \
  "mov.b64 da, {al, %3};\n\tmov.b64 db, {bl, %3};\n\t"                      \
  "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, p;\n\t"
template <int NK16>
__device__ __forceinline__ void mma_ss(uint32_t tacc, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t ak16,
                                       uint32_t bk16, uint32_t idesc, uint32_t acc) {
  static_assert(NK16 == 1 || NK16 == 2 || NK16 == 4, "mma_ss: unsupported block");
  if (NK16 == 1)
    asm volatile(LUSH_SS_FIRST "}" ::"r"(tacc), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(ak16), "r"(bk16), "r"(idesc),
                 "r"(acc)
                 : "memory");
  else if (NK16 == 2)
    asm volatile(LUSH_SS_FIRST LUSH_SS_NEXT "}" ::"r"(tacc), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(ak16), "r"(bk16),
                 "r"(idesc), "r"(acc)
                 : "memory");
  else
    asm volatile(LUSH_SS_FIRST LUSH_SS_NEXT LUSH_SS_NEXT LUSH_SS_NEXT "}" ::"r"(tacc), "r"(a_lo), "r"(b_lo), "r"(hi),
                 "r"(ak16), "r"(bk16), "r"(idesc), "r"(acc)
                 : "memory");
}
#undef LUSH_SS_FIRST
#undef LUSH_SS_NEXT

// same with the A operand in tensor memory: 8 columns (16 fp16) of all 128 lanes per K=16.  The 64 K-values of one
// activation group sit where the two epilogue warpgroups found their accumulator chunks: K [0,32) in columns [0,16) of the
// group's 64-column slot, K [32,64) in columns [32,48)
#define LUSH_TS_FIRST                                                       \
  "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t.reg .b32 at, bl;\n\t"            \
  "setp.ne.b32 p, %6, 0;\n\t"                                               \
  "mov.b32 at, %1;\n\tmov.b32 bl, %2;\n\t"                                  \
  "mov.b64 db, {bl, %3};\n\t"                                              \
  "tcgen05.mma.cta_group::1.kind::f16 [%0], [at], db, %5, p;\n\t"          \
  "setp.eq.b32 p, %6, %6;\n\t"
#define LUSH_TS_NEXT(STEP)                                                  \
  "add.u32 at, at, " #STEP ";\n\tadd.u32 bl, bl, %4;\n\t"                   \
  "mov.b64 db, {bl, %3};\n\t"                                              \
  "tcgen05.mma.cta_group::1.kind::f16 [%0], [at], db, %5, p;\n\t"
__device__ __forceinline__ void mma_ts4(uint32_t tacc, uint32_t a_tmem, uint32_t b_lo, uint32_t hi, uint32_t bk16,
                                        uint32_t idesc, uint32_t acc) {
  asm volatile(LUSH_TS_FIRST LUSH_TS_NEXT(8) LUSH_TS_NEXT(24) LUSH_TS_NEXT(8) "}" ::"r"(tacc), "r"(a_tmem), "r"(b_lo), "r"(hi),
               "r"(bk16), "r"(idesc), "r"(acc)
               : "memory");
}
#undef LUSH_TS_FIRST
#undef LUSH_TS_NEXT

// one lane of a converged warp (the same lane every time for a full warp)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- warps 14 and 15 (all 32 lanes, converged): issuer `my` issues blocks my, my+2, ... of every step.
// Both walk the same block sequence of the weight ring and skip each other's blocks.
template <class R>
__device__ __forceinline__ void chain_mma_issuer(ChainBars& B, uint8_t* smem, uint32_t tmem_base, const ChainPlan& P,
                                                 const ChainMap map, int64_t ntiles, const uint32_t my) {
  uint32_t stage = 0, phase = 0;
  uint32_t a_par = 0;                    // parity of the next a_ready phase (all four groups advance together)
  uint32_t n_used0 = 0, n_used1 = 0;     // how often each accumulator buffer has been written so far
  uint32_t tok_par = 0, acc_sel = 0;      // tok_par: parity of the next unobserved phase of the OTHER issuer's token
  const uint32_t sbase = smem_u32(smem);
  const uint32_t hi128 = desc_hi(128u);
  constexpr uint32_t A_K16 = (2u * TROWS * 16u) >> 4;   // start-address units per 16 columns of a 128-row blob
  const uint32_t act_lo = desc_lo(sbase + R::ACT, TROWS * 16u);
  const uint32_t enc_lo = desc_lo(sbase + map.enc, TROWS * 16u);
  int64_t it = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int b = (int)(it & 1);
    if (my == 0) mbar_wait(&B.pro_ready[b], (uint32_t)((it >> 1) & 1));   // only block 0 reads prologue operands
    const uint32_t encv_lo = desc_lo(sbase + map.encv + b * map.encv_stride, TROWS * 16u);
    const uint32_t x16_lo = desc_lo(sbase + map.x16 + b * map.x16_stride, TROWS * 16u);
    for (int s = 0; s < P.nsteps; ++s) {
      const ChainStep S = P.steps[s];
      const uint32_t idesc = make_idesc_f16(TROWS, S.N, 0, 0);
      const uint32_t b_lbo = (uint32_t)S.N * 16u;
      const uint32_t B_K16 = (2u * b_lbo) >> 4;
      const uint32_t e_lo = S.enc_src ? encv_lo : enc_lo;
      const uint32_t merge = (uint32_t)chain_merge_h0(S, R::STAGE_BYTES);
      const uint32_t nblk = (uint32_t)chain_nblocks(S, R::STAGE_BYTES);
      const uint32_t tacc = tmem_base + acc_sel * 256u;
      uint32_t seen = 0;                 // activation groups whose completion this issuer has observed this step
      chain_trace(P, it, s, 2 * (int)my);
      for (uint32_t i = 0; i < nblk; ++i) {
        if ((i & 1u) != my) {                  // the other issuer's block
          if (++stage == R::NSTAGE) { stage = 0; phase ^= 1u; }
          continue;
        }
        if (i == 0) {                          // overwrites accumulator acc_sel: its previous epilogue must have drained it
          const uint32_t used = acc_sel ? n_used1 : n_used0;
          if (used) mbar_wait(&B.d_free[acc_sel], (used - 1u) & 1u);
        }
        if (i > 0) {                           // must not overtake block i-1 (the other issuer's) in the tensor pipe:
          mbar_wait(&B.issued[my ^ 1u], tok_par);   // fp32 accumulation order = block order, run after run
          tok_par ^= 1u;
        }
        const bool last_blk = (i + 1u == nblk);                // probes 12..15: the step's last block
        if (last_blk) chain_trace(P, it, s, 12);
        mbar_wait(&B.full[stage], phase);
        if (last_blk) chain_trace(P, it, s, 13);
        uint32_t b_lo = desc_lo(sbase + R::STAGE0 + stage * R::STAGE_BYTES, b_lbo);
        const uint32_t grp = (i == 0) ? 0u : i - 1u + merge;   // activation group read by this block (if any)
        const bool has_h = (i > 0) || merge;
        if (has_h && S.wait_a) { mbar_wait(&B.a_ready[grp], a_par); seen |= (1u << grp); }
        if (last_blk) chain_trace(P, it, s, 14);
        tc_fence_after();
        if (elect_one()) {
          uint32_t acc = (i == 0) ? 0u : 1u;
          if (i == 0) {
            if (S.pre16) { mma_ss<1>(tacc, x16_lo, b_lo, hi128, A_K16, B_K16, idesc, acc); acc = 1; b_lo += B_K16; }
            if (S.enc_slices == 2) { mma_ss<4>(tacc, e_lo, b_lo, hi128, A_K16, B_K16, idesc, acc); acc = 1; b_lo += 4u * B_K16; }
            else if (S.enc_slices == 1) { mma_ss<2>(tacc, e_lo, b_lo, hi128, A_K16, B_K16, idesc, acc); acc = 1; b_lo += 2u * B_K16; }
          }
          if (has_h) {
            if (R::ACT_IN_TMEM)   // A = packed activations inside the OTHER accumulator buffer, group grp in its slot
              mma_ts4(tacc, tmem_base + (acc_sel ^ 1u) * 256u + grp * 64u, b_lo, hi128, B_K16, idesc, acc);
            else
              mma_ss<4>(tacc, act_lo + grp * 4u * A_K16, b_lo, hi128, A_K16, B_K16, idesc, acc);
          }
          umma_commit(&B.empty[stage]);
          mbar_arrive(&B.issued[my]);
          if (i == 0) {
            if (S.free_enc == 1) umma_commit(&B.enc_free);
            else if (S.free_enc == 2) umma_commit(&B.pro_free[b]);
          }
        }
        if (last_blk) chain_trace(P, it, s, 15);
        if (++stage == R::NSTAGE) { stage = 0; phase ^= 1u; }
      }
      // per-step bookkeeping (both issuers observe every barrier phase they track, in order)
      if (S.wait_a) {   // every writer publishes all four groups (a 128-wide one publishes 2 and 3 as empty)
        for (uint32_t j = 0; j < 4u; ++j)
          if (!((seen >> j) & 1u)) mbar_wait(&B.a_ready[j], a_par);
        a_par ^= 1u;
      }
      // the other issuer's last block of this step has no successor of mine inside the step when it is the step's last
      // block: observe its token here (every phase of a barrier is observed exactly once by its one waiter)
      if (((nblk & 1u) != 0u) == (my == 1u)) {
        mbar_wait(&B.issued[my ^ 1u], tok_par);
        tok_par ^= 1u;
      }
      if (elect_one()) umma_commit(&B.d_ready[acc_sel]);
      if (acc_sel) ++n_used1; else ++n_used0;
      if (!S.keep_acc) acc_sel ^= 1u;
      chain_trace(P, it, s, 2 * (int)my + 1);
    }
  }
}

// epilogue side of the hand-off: this warp's st.shared into columns [64j, 64j+64) of the activation tile are done
__device__ __forceinline__ void chain_publish(ChainBars& B, int j, int lane) {
  fence_proxy_async();
  __syncwarp();
  mbar_arrive_if(&B.a_ready[j], lane == 0);
}
// the same hand-off when the activations went to tensor memory (tcgen05.st): no proxy fence, the stores are ordered
// before the arrive by wait::st + the tcgen05 fence
__device__ __forceinline__ void chain_publish_tmem(ChainBars& B, int j, int lane) {
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  mbar_arrive_if(&B.a_ready[j], lane == 0);
}
// this warp's share of accumulator d is in registers
__device__ __forceinline__ void chain_release_acc(ChainBars& B, int d, int lane) {
  tc_fence_before();
  __syncwarp();
  mbar_arrive_if(&B.d_free[d], lane == 0);
}

}  // namespace lush
