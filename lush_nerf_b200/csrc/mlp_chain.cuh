// Skeleton shared by the fused tcgen05 MLP kernels (forward: mlp_tc.cu, dgrad: mlp_tc_bwd.cu).
//
// A "chain" is a sequence of GEMM steps  D_s = X_s . W_s^T  over ONE 128-row sample tile whose intermediate
// activations never leave the SM:
//   * accumulators live in TMEM columns [0,256): two regions of 128 columns, one per N-half of a layer;
//   * the fp16 activations that feed the next step live in TMEM too (columns [256,512): two buffers of 128
//     columns = 256 K-values each, ping-pong by step) and are consumed as the A operand straight from tensor
//     memory (tcgen05.mma [d], [a], b-desc);  encodings / bias-ones / head-gradient tiles are small SS-mode
//     operands in shared memory;
//   * shared memory is almost entirely a ring (5 x 32 KB) of weight blocks [128 out x 128 K] streamed with
//     bulk async copies (TMA engine) in exactly the order the MMAs consume them.
// Each layer is issued as two N=128 halves, strictly in order (s,h0) (s,h1) (s+1,h0) ...  While the tensor core
// works on half 1 of layer s, ALL eight epilogue warps drain half 0 (tcgen05.ld -> bias/activation -> fp16 ->
// tcgen05.st; warpgroup g takes 64 of the 128 columns) and publish it as the first 128 K-values of the next
// layer's A operand (one mbarrier per half), so layer s+1 starts on them while half 1 is being drained: the
// tensor pipe never waits for a whole-layer hand-off.
//
// Warp roles (512 threads, 1 CTA per SM, persistent over tiles):
//   warps 0-3     prologue warpgroup: builds the NEXT tile's shared-memory operands (positional encodings /
//                 head-gradient tile) while the current tile is in flight (double-buffered)
//   warps 4-7     epilogue warpgroup 0  (columns  0..63  of every output half)
//   warps 8-11    epilogue warpgroup 1  (columns 64..127 of every output half)
//   warp 12       weight producer (one lane)
//   warps 14, 15  MMA issuers (converged warps, one elected lane issues): warp 1+r issues everything that
//                 targets accumulator region r (single-thread issue costs ~80 cycles per MMA, more than the
//                 64 cycles an N=128 MMA occupies the tensor pipe, so the work is split); a token barrier keeps
//                 the pipe order (s,h0) before (s,h1).  Warp 14 also owns the TMEM allocation.
// The SM's warp arbiter favours the highest warp id among eligible warps: the latency-critical single-thread
// issuers sit at the top, the ALU-heavy but latency-tolerant encoder warps at the bottom.
#pragma once
#include <stdlib.h>
#include "mlp_tc.cuh"

namespace lush {

constexpr int CH_THREADS = 512;
constexpr int CH_NSTAGE = 5;
constexpr int CH_STAGE_BYTES = 128 * 128 * 2;  // 32 KB: one [128 x 128] fp16 weight block (K = 128)
constexpr int CH_MAX_STEPS = kMaxDepth + 6;
constexpr int CH_WG_PRO = 0;                   // first warp of the prologue warpgroup
constexpr int CH_WG_EPI0 = 4;                  // first warp of epilogue warpgroup 0
constexpr int CH_WARP_PRODUCER = 12;
constexpr int CH_WARP_MMA0 = 14;               // issuer of accumulator region 0 (region 1: CH_WARP_MMA0 + 1)
constexpr int CH_X16_BYTES = TROWS * 16 * 2;   // [128 x 16] fp16 tile

struct ChSmem {
  static constexpr int STAGE0 = 0;
  static constexpr int ENC0 = STAGE0 + CH_NSTAGE * CH_STAGE_BYTES;   // 2 x [128 x 64]
  static constexpr int ENCV0 = ENC0 + 2 * ENC_BYTES;                 // 2 x [128 x 32]
  static constexpr int X160 = ENCV0 + 2 * ENCV_BYTES;                // 2 x [128 x 16]
  static constexpr int PARAMS = X160 + 2 * CH_X16_BYTES;
};

// One GEMM step.  Its K range is, in this order: an optional K=16 slice against the X16 shared-memory tile
// (bias ones / head gradients), optional K=32 slices against an encoding tile in shared memory, and h_slices
// K=32 slices against the activation buffer in tensor memory (h_slices is a multiple of 4).
// Weight stream per output half: one "small" block [pre16 | enc] (if any) followed by h_slices/4 blocks
// [NH x 128]; every block is one ring stage, each sub-block in the tile-canonical layout with NH rows.
struct ChainStep {
  int8_t nhalf;       // 1 | 2 output halves
  int8_t d0;          // accumulator region (= epilogue warpgroup) of half 0; half 1 always uses region 1
  int16_t NH;         // output columns per half (128 | 64)
  int8_t pre16;       // 0 | 1
  int8_t enc_slices;  // 0 | 1 | 2
  int8_t enc_src;     // 0: ENC (64 columns), 1: ENCV (32 columns)
  int8_t h_slices;    // 0 | 4 | 8
  int8_t a_in;        // activation buffer read by the TS slices (0 | 1)
  int8_t wait_a;      // 1: that buffer is being written by the previous step's epilogue (wait per 64 columns)
  int8_t a_out;       // buffer the epilogue writes (-1: none)
  int8_t epi;         // kernel-specific epilogue kind
  int16_t p0, p1;     // epilogue parameters
  int32_t off0, off1; // epilogue byte offsets
};

struct ChainPlan {
  int nsteps;
  int debug;              // experiments only (LUSH_CHAIN_DEBUG): bit 0 = do not stream weights (results are garbage)
  int64_t stream_bytes;   // weight stream per tile
  ChainStep steps[CH_MAX_STEPS];
};

__host__ __device__ inline int chain_small_bytes(const ChainStep& s) {
  return s.NH * 2 * (16 * s.pre16 + KS * s.enc_slices);
}
__host__ __device__ inline int chain_half_bytes(const ChainStep& s) {
  return chain_small_bytes(s) + s.NH * 2 * KS * s.h_slices;
}
static inline void chain_finish_plan(ChainPlan& P) {
  const char* dbg = getenv("LUSH_CHAIN_DEBUG");
  P.debug = dbg ? atoi(dbg) : 0;
  int64_t b = 0;
  for (int s = 0; s < P.nsteps; ++s) b += (int64_t)P.steps[s].nhalf * chain_half_bytes(P.steps[s]);
  P.stream_bytes = b;
}

// ---- optional timeline trace (LUSH_CHAIN_DEBUG bit 1): CTA 0 records clock64() of its third tile's events
//   slot = step*16 + {0,1: issue of half 0 start/end; 2,3: half 1; 4 + 4*half + {0: d_ready seen, 1: accumulator
//   released, 2: published} for epilogue warpgroup 0; 12..: warpgroup 1 half 0 d_ready seen / published}
static __device__ long long g_chain_trace[CH_MAX_STEPS * 16];   // one copy per translation unit
__device__ __forceinline__ void chain_trace(const ChainPlan& P, int64_t it, int step, int ev) {
  if ((P.debug & 2) && blockIdx.x == 0 && it == 2 && (threadIdx.x & 31) == 0) g_chain_trace[step * 16 + ev] = clock64();
}

struct ChainBars {
  uint64_t full[CH_NSTAGE], empty[CH_NSTAGE];
  uint64_t a_ready[2][2][2];    // [issuer][buf][j]: K-values [128j, 128j+128) of activation buffer `buf` are complete
                                // (8 warps).  One set per issuer (each tracks its own parity) and per buffer: a
                                // barrier then completes again only two steps later, which transitively requires
                                // both issuers to have observed the previous completion (no parity aliasing)
  uint64_t h0_issued;           // issuer 0 has issued half 0 of the current two-half step (pipe order token)
  uint64_t d_ready[2];          // accumulator region r complete (MMA commit)
  uint64_t d_free[2];           // accumulator region r is in the epilogue warps' registers (8 warps)
  uint64_t pro_ready[2];        // prologue operands of tile parity b are in shared memory (4 warps)
  uint64_t pro_free[2];         // ... and no longer read by any MMA (commit)
  uint64_t aux;                 // kernel-specific (forward: head partial sums of warpgroup 0, 4 warps)
};

__device__ __forceinline__ void chain_init_bars(ChainBars& B) {
  for (int s = 0; s < CH_NSTAGE; ++s) { mbar_init(&B.full[s], 1); mbar_init(&B.empty[s], 1); }
  for (int r = 0; r < 2; ++r) {
    for (int i = 0; i < 2; ++i) { mbar_init(&B.a_ready[i][r][0], 8); mbar_init(&B.a_ready[i][r][1], 8); }
    mbar_init(&B.d_ready[r], 1); mbar_init(&B.d_free[r], 8);
    mbar_init(&B.pro_ready[r], 4); mbar_init(&B.pro_free[r], 2);
  }
  mbar_init(&B.h0_issued, 1);
  mbar_init(&B.aux, 4);
  fence_barrier_init();
}

// ---- warp 0, one lane: stream the packed weight blocks of every tile through the ring
__device__ __forceinline__ void chain_producer(ChainBars& B, uint8_t* smem, const uint8_t* __restrict__ wstream,
                                               const ChainPlan& P, int64_t ntiles) {
  int stage = 0; uint32_t phase = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const uint8_t* src = wstream;
    for (int s = 0; s < P.nsteps; ++s) {
      const ChainStep& S = P.steps[s];
      const uint32_t small = (uint32_t)chain_small_bytes(S), big = (uint32_t)S.NH * 128 * 2;
      const int nbig = S.h_slices >> 2;
      for (int h = 0; h < S.nhalf; ++h) {
        for (int j = (small ? -1 : 0); j < nbig; ++j) {
          const uint32_t bytes = j < 0 ? small : big;
          mbar_wait(&B.empty[stage], phase ^ 1);
          if (P.debug & 1) { mbar_arrive(&B.full[stage]); }
          else {
            mbar_arrive_expect_tx(&B.full[stage], bytes);
            bulk_g2s(smem + ChSmem::STAGE0 + stage * CH_STAGE_BYTES, src, bytes, &B.full[stage]);
          }
          src += bytes;
          if (++stage == CH_NSTAGE) { stage = 0; phase ^= 1; }
        }
      }
    }
  }
}

// ---- MMA issue blocks.  The whole issuer warp executes them converged; only the elected lane issues.
#define LUSH_TS_FIRST                                                       \
  "{\n\t.reg .pred q, p;\n\t.reg .b64 db;\n\t.reg .b32 a, bl;\n\t"         \
  "setp.ne.b32 p, %6, 0;\n\t"                                               \
  "mov.b32 a, %1;\n\tmov.b32 bl, %2;\n\t"                                   \
  "mov.b64 db, {bl, %3};\n\t"                                               \
  "tcgen05.mma.cta_group::1.kind::f16 [%0], [a], db, %5, p;\n\t"           \
  "setp.eq.b32 p, %6, %6;\n\t"
#define LUSH_TS_NEXT                                                        \
  "add.u32 a, a, 8;\n\tadd.u32 bl, bl, %4;\n\t"                             \
  "mov.b64 db, {bl, %3};\n\t"                                               \
  "tcgen05.mma.cta_group::1.kind::f16 [%0], [a], db, %5, p;\n\t"
// 4 consecutive K=16 MMAs (64 K-values): A from tensor memory (8 columns per K=16), B advancing by bk16
__device__ __forceinline__ void mma_ts_k64(uint32_t tacc, uint32_t a_tmem, uint32_t b_lo,
                                           uint32_t b_hi, uint32_t bk16, uint32_t idesc, uint32_t acc) {
  asm volatile(LUSH_TS_FIRST LUSH_TS_NEXT LUSH_TS_NEXT LUSH_TS_NEXT "}" ::"r"(tacc), "r"(a_tmem), "r"(b_lo),
               "r"(b_hi), "r"(bk16), "r"(idesc), "r"(acc)
               : "memory");
}
#undef LUSH_TS_FIRST
#undef LUSH_TS_NEXT

#define LUSH_SS_FIRST                                                       \
  "{\n\t.reg .pred q, p;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"    \
  "setp.ne.b32 p, %7, 0;\n\t"                                               \
  "mov.b32 al, %1;\n\tmov.b32 bl, %2;\n\t"                                  \
  "mov.b64 da, {al, %3};\n\tmov.b64 db, {bl, %3};\n\t"                      \
  "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, p;\n\t"            \
  "setp.eq.b32 p, %7, %7;\n\t"
#define LUSH_SS_NEXT                                                        \
  "add.u32 al, al, %4;\n\tadd.u32 bl, bl, %5;\n\t"                          \
  "mov.b64 da, {al, %3};\n\tmov.b64 db, {bl, %3};\n\t"                      \
  "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %6, p;\n\t"
// NK16 consecutive K=16 MMAs with both operands in shared memory (descriptor low words a_lo / b_lo advance by
// ak16 / bk16 per K=16; both share the high word `hi`)
template <int NK16>
__device__ __forceinline__ void mma_ss(uint32_t tacc, uint32_t a_lo, uint32_t b_lo, uint32_t hi,
                                       uint32_t ak16, uint32_t bk16, uint32_t idesc, uint32_t acc) {
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

// one lane of a converged warp (the same lane every time for a full warp)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- warps 1 and 2 (all 32 lanes, converged): issuer `my` issues every MMA that targets accumulator region
// `my`.  Both walk the same block sequence of the weight ring and skip each other's blocks.
__device__ __forceinline__ void chain_mma_issuer(ChainBars& B, uint8_t* smem, uint32_t tmem_base,
                                                 const ChainPlan& P, int64_t ntiles, const uint32_t my) {
  uint32_t stage = 0, phase = 0;
  uint32_t a_phase = 0, dfree_phase = 0, d_used = 0, tok_phase = 0;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t hi128 = desc_hi(128u);
  const uint32_t tacc = tmem_base + my * 128u;
  constexpr uint32_t A_K16 = (2u * TROWS * 16u) >> 4;   // start-address units per 16 columns of a 128-row blob
  int64_t it = 0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int b = (int)(it & 1);
    mbar_wait(&B.pro_ready[b], (uint32_t)((it >> 1) & 1));
    const uint32_t enc_lo = desc_lo(sbase + ChSmem::ENC0 + b * ENC_BYTES, TROWS * 16u);
    const uint32_t encv_lo = desc_lo(sbase + ChSmem::ENCV0 + b * ENCV_BYTES, TROWS * 16u);
    const uint32_t x16_lo = desc_lo(sbase + ChSmem::X160 + b * CH_X16_BYTES, TROWS * 16u);
    for (int s = 0; s < P.nsteps; ++s) {
      const ChainStep S = P.steps[s];
      const uint32_t idesc = make_idesc_f16(TROWS, S.NH, 0, 0);
      const uint32_t b_lbo = (uint32_t)S.NH * 16u;
      const uint32_t B_K16 = (2u * b_lbo) >> 4;
      const uint32_t a_tmem = tmem_base + 256u + (uint32_t)S.a_in * 128u;
      const uint32_t e_lo = S.enc_src ? encv_lo : enc_lo;
      const uint32_t has_small = (S.pre16 || S.enc_slices) ? 1u : 0u;
      const uint32_t nbig = (uint32_t)S.h_slices >> 2;
      const uint32_t a_todo = S.wait_a ? nbig : 0u;   // 128-K groups of the new A buffer this issuer must observe
      uint32_t a_seen = 0;
      for (int h = 0; h < S.nhalf; ++h) {
        const uint32_t reg = (h == 0) ? (uint32_t)S.d0 : 1u;
        if (reg != my) {                       // the other issuer's blocks
          stage += has_small + nbig;
          if (stage >= CH_NSTAGE) { stage -= CH_NSTAGE; phase ^= 1u; }
          continue;
        }
        if (d_used) { mbar_wait(&B.d_free[my], dfree_phase); dfree_phase ^= 1u; }
        d_used = 1u;
        if (h == 1 && !(P.debug & 16)) { mbar_wait(&B.h0_issued, tok_phase); tok_phase ^= 1u; }   // keep the pipe order (s,h0) (s,h1)
        chain_trace(P, it, s, 2 * h);
        uint32_t acc = 0;
        if (has_small) {
          mbar_wait(&B.full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            uint32_t b_lo = desc_lo(sbase + ChSmem::STAGE0 + stage * CH_STAGE_BYTES, b_lbo);
            uint32_t acc1 = 0;
            if (S.pre16) { mma_ss<1>(tacc, x16_lo, b_lo, hi128, A_K16, B_K16, idesc, acc1); acc1 = 1; b_lo += B_K16; }
            if (S.enc_slices == 2) mma_ss<4>(tacc, e_lo, b_lo, hi128, A_K16, B_K16, idesc, acc1);
            else if (S.enc_slices == 1) mma_ss<2>(tacc, e_lo, b_lo, hi128, A_K16, B_K16, idesc, acc1);
            umma_commit(&B.empty[stage]);
          }
          acc = 1;
          if (++stage == CH_NSTAGE) { stage = 0; phase ^= 1u; }
        }
        for (uint32_t j = 0; j < nbig; ++j) {
          mbar_wait(&B.full[stage], phase);
          if (a_seen < a_todo) {
            const uint32_t k = 2u * (uint32_t)S.a_in + j;
            mbar_wait(&B.a_ready[my][S.a_in][j], (a_phase >> k) & 1u); a_phase ^= (1u << k); a_seen = j + 1;
          }
          tc_fence_after();
          if (elect_one()) {
            const uint32_t b_lo = desc_lo(sbase + ChSmem::STAGE0 + stage * CH_STAGE_BYTES, b_lbo);
            mma_ts_k64(tacc, a_tmem + j * 64u, b_lo, hi128, B_K16, idesc, acc);
            mma_ts_k64(tacc, a_tmem + j * 64u + 32u, b_lo + 4u * B_K16, hi128, B_K16, idesc, 1u);
            umma_commit(&B.empty[stage]);
          }
          acc = 1;
          if (j == 0) chain_trace(P, it, s, 14 + h);
          if (++stage == CH_NSTAGE) { stage = 0; phase ^= 1u; }
        }
        if (elect_one()) {
          umma_commit(&B.d_ready[my]);
          if (h == 0 && S.nhalf == 2) mbar_arrive(&B.h0_issued);
        }
        chain_trace(P, it, s, 2 * h + 1);
      }
      // groups this issuer did not need (it issued nothing in this step): observe them anyway so that the
      // parity bookkeeping of a_ready[my][*] never skips a phase
      for (; a_seen < a_todo; ++a_seen) {
        const uint32_t k = 2u * (uint32_t)S.a_in + a_seen;
        mbar_wait(&B.a_ready[my][S.a_in][a_seen], (a_phase >> k) & 1u); a_phase ^= (1u << k);
      }
    }
    if (elect_one()) umma_commit(&B.pro_free[b]);   // my MMAs reading this tile's shared-memory operands are complete
  }
}

// epilogue side of the hand-off: this warp's tcgen05.st into K-values [128j, 128j+128) of buffer `buf` are complete
__device__ __forceinline__ void chain_publish(ChainBars& B, int buf, int j, int lane) {
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) { mbar_arrive(&B.a_ready[0][buf][j]); mbar_arrive(&B.a_ready[1][buf][j]); }
}
// this warp's share of accumulator region r is in registers
__device__ __forceinline__ void chain_release_acc(ChainBars& B, int r, int lane) {
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(&B.d_free[r]);
}

}  // namespace lush
