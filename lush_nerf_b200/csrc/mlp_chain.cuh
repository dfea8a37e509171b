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
// Each layer is issued as two N=128 halves.  While the tensor core works on half 1 of layer s, epilogue
// warpgroup 0 drains half 0 (tcgen05.ld -> activation -> fp16 -> tcgen05.st) and publishes the next layer's
// A operand 64 columns at a time (one mbarrier each), so layer s+1 starts as soon as its first K-slices
// exist: the tensor pipe does not wait for a whole-layer hand-off.
//
// Warp roles (512 threads, 1 CTA per SM, persistent over tiles):
//   warp 0        weight producer (one lane)
//   warp 1        MMA issuer (converged warp, elected lane issues), TMEM allocation
//   warps 4-7     epilogue warpgroup 0  (accumulator region 0)
//   warps 8-11    epilogue warpgroup 1  (accumulator region 1)
//   warps 12-15   prologue warpgroup: builds the NEXT tile's shared-memory operands (positional encodings /
//                 head-gradient tile) while the current tile is in flight (double-buffered)
#pragma once
#include "mlp_tc.cuh"

namespace lush {

constexpr int CH_THREADS = 512;
constexpr int CH_NSTAGE = 5;
constexpr int CH_STAGE_BYTES = 128 * 128 * 2;  // 32 KB: one [128 x 128] fp16 weight block (K = 128)
constexpr int CH_MAX_STEPS = kMaxDepth + 6;
constexpr int CH_WG_EPI0 = 4;                  // first warp of epilogue warpgroup 0
constexpr int CH_WG_PRO = 12;                  // first warp of the prologue warpgroup
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
  int64_t b = 0;
  for (int s = 0; s < P.nsteps; ++s) b += (int64_t)P.steps[s].nhalf * chain_half_bytes(P.steps[s]);
  P.stream_bytes = b;
}

struct ChainBars {
  uint64_t full[CH_NSTAGE], empty[CH_NSTAGE];
  uint64_t a_ready[4];          // 64-column group j of the activation buffer being written is complete (4 warps)
  uint64_t d_ready[2];          // accumulator region r complete (MMA commit)
  uint64_t d_free[2];           // accumulator region r drained by its epilogue warpgroup (4 warps)
  uint64_t pro_ready[2];        // prologue operands of tile parity b are in shared memory (4 warps)
  uint64_t pro_free[2];         // ... and no longer read by any MMA (commit)
  uint64_t aux;                 // kernel-specific (forward: sigma partial sums of warpgroup 0)
};

__device__ __forceinline__ void chain_init_bars(ChainBars& B) {
  for (int s = 0; s < CH_NSTAGE; ++s) { mbar_init(&B.full[s], 1); mbar_init(&B.empty[s], 1); }
  for (int j = 0; j < 4; ++j) mbar_init(&B.a_ready[j], 4);
  for (int r = 0; r < 2; ++r) {
    mbar_init(&B.d_ready[r], 1); mbar_init(&B.d_free[r], 4);
    mbar_init(&B.pro_ready[r], 4); mbar_init(&B.pro_free[r], 1);
  }
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
          mbar_arrive_expect_tx(&B.full[stage], bytes);
          bulk_g2s(smem + ChSmem::STAGE0 + stage * CH_STAGE_BYTES, src, bytes, &B.full[stage]);
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

// ---- warp 1 (all 32 lanes, converged): issue every MMA of every tile
__device__ __forceinline__ void chain_mma_issuer(ChainBars& B, uint8_t* smem, uint32_t tmem_base,
                                                 const ChainPlan& P, int64_t ntiles, int lane) {
  int stage = 0; uint32_t phase = 0;
  uint32_t a_phase = 0;                 // bit j: parity to wait for on a_ready[j]
  uint32_t dfree_phase = 0, d_used = 0; // bit r
  const uint32_t sbase = smem_u32(smem);
  const uint32_t hi128 = desc_hi(128u);
  constexpr uint32_t A_K16 = (2u * TROWS * 16u) >> 4;   // start-address units per 16 columns of a 128-row blob
  int64_t it = 0;
  // the full barrier of the next stage is polled one block ahead so its latency hides behind the MMA issue
  bool full_ok = mbar_test_wait(&B.full[0], 0);
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
      const bool has_small = S.pre16 || S.enc_slices;
      const int nbig = S.h_slices >> 2;
      for (int h = 0; h < S.nhalf; ++h) {
        const uint32_t reg = (h == 0) ? (uint32_t)S.d0 : 1u;
        if ((d_used >> reg) & 1u) {
          mbar_wait(&B.d_free[reg], (dfree_phase >> reg) & 1u); dfree_phase ^= (1u << reg);
        }
        d_used |= (1u << reg);
        const uint32_t tacc = tmem_base + reg * 128u;
        uint32_t acc = 0;
        for (int j = (has_small ? -1 : 0); j < nbig; ++j) {
          if (!full_ok) mbar_wait(&B.full[stage], phase);
          const int nstage = (stage + 1 == CH_NSTAGE) ? 0 : stage + 1;
          const uint32_t nphase = (stage + 1 == CH_NSTAGE) ? (phase ^ 1u) : phase;
          full_ok = mbar_test_wait(&B.full[nstage], nphase);
          uint32_t b_lo = desc_lo(sbase + ChSmem::STAGE0 + stage * CH_STAGE_BYTES, b_lbo);
          if (j < 0) {
            tc_fence_after();
            if (elect_one()) {
              uint32_t acc1 = acc;
              if (S.pre16) {
                mma_ss<1>(tacc, x16_lo, b_lo, hi128, A_K16, B_K16, idesc, acc1); acc1 = 1;
                b_lo += B_K16;
              }
              if (S.enc_slices == 2) mma_ss<4>(tacc, e_lo, b_lo, hi128, A_K16, B_K16, idesc, acc1);
              else if (S.enc_slices == 1) mma_ss<2>(tacc, e_lo, b_lo, hi128, A_K16, B_K16, idesc, acc1);
              umma_commit(&B.empty[stage]);
            }
            __syncwarp();
            acc = 1;
          } else {
            const bool wa = (h == 0) && S.wait_a;
            if (wa) { mbar_wait(&B.a_ready[2 * j], (a_phase >> (2 * j)) & 1u); a_phase ^= (1u << (2 * j)); }
            tc_fence_after();
            if (elect_one()) mma_ts_k64(tacc, a_tmem + (uint32_t)j * 64u, b_lo, hi128, B_K16, idesc, acc);
            __syncwarp();
            acc = 1;
            if (wa) {
              mbar_wait(&B.a_ready[2 * j + 1], (a_phase >> (2 * j + 1)) & 1u); a_phase ^= (1u << (2 * j + 1));
              tc_fence_after();
            }
            if (elect_one()) {
              mma_ts_k64(tacc, a_tmem + (uint32_t)j * 64u + 32u, b_lo + 4u * B_K16, hi128, B_K16, idesc, 1u);
              umma_commit(&B.empty[stage]);
            }
            __syncwarp();
          }
          stage = nstage; phase = nphase;
        }
        if (elect_one()) umma_commit(&B.d_ready[reg]);
        __syncwarp();
      }
    }
    if (elect_one()) umma_commit(&B.pro_free[b]);   // every MMA reading this tile's shared-memory operands is complete
    __syncwarp();
  }
}

// epilogue side of the hand-off: this warp's tcgen05.st of 64-column group j are complete -> publish
__device__ __forceinline__ void chain_publish(ChainBars& B, int j, int lane) {
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(&B.a_ready[j]);
}
// all accumulator columns of region r are in registers
__device__ __forceinline__ void chain_release_acc(ChainBars& B, int r, int lane) {
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(&B.d_free[r]);
}

}  // namespace lush
