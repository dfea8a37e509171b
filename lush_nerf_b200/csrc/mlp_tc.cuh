// Shared plan / layout definitions of the tcgen05 MLP kernels (forward: mlp_tc.cu, backward: mlp_tc_bwd.cu).
#pragma once
#include "mlp_common.cuh"
#include "tc_common.cuh"

namespace lush {
using namespace tc;

constexpr int TW = 256;            // network width this path is built for
constexpr int TROWS = 128;         // rows per tile (= TMEM lanes)
constexpr int KS = 32;             // K elements per weight slice
constexpr int NSTAGE = 3;
constexpr int STAGE_BYTES = TW * KS * 2;          // 16 KB
constexpr int H_BYTES = TROWS * TW * 2;           // 64 KB
constexpr int ENC_BYTES = TROWS * 64 * 2;         // 16 KB
constexpr int ENCV_BYTES = TROWS * 32 * 2;        // 8 KB
constexpr int HV_BYTES = TROWS * 128 * 2;         // 32 KB
constexpr int MAX_STEPS = kMaxDepth + 2;
constexpr int FWD_THREADS = 352;   // warp 0 producer, warps 1-2 MMA issuers (one per tile), warps 3-10 epilogue
constexpr int EPI_WARP0 = 3;

enum FwdEpi { EPI_TRUNK = 0, EPI_FEATURE = 1, EPI_VIEWS = 2 };

struct FwdStep {
  int N;           // output width (256 or 128)
  int nslices;     // K slices of 32
  int enc_slices;  // leading slices read the encoding buffer (ENC), the rest read H
  int enc_first;   // 1: encoding segment comes first (layers 0 / skip+1); 0: H first, encoding last (views)
  int epi;         // FwdEpi
  int bias_src;    // 0: pts_linears[i].bias, 1: feature_linear.bias, 2: views_linears.0.bias (folded into the MMA)
  int last_trunk;  // 1: also produce sigma = h . w_alpha + b
  int stash_slot;  // index of the H blob this step's OUTPUT occupies in the stash record (training)
};

struct FwdPlan {
  int nsteps;
  int D;
  FwdStep steps[MAX_STEPS];
  // smem parameter block (floats): w_alpha[256], w_rgb[3][128], b_alpha, b_rgb[3]
  // (layer biases ride in the weight stream as one extra K=16 slice per step against a ones column)
  int w_alpha_off, w_rgb_off, b_head_off, bias_off, nparam_floats;   // bias_off: step s -> [bias_off + s*256, +N)
  // stash record (bytes per tile)
  int rec_bytes, rec_encp, rec_h0, rec_feat, rec_encv, rec_hv, rec_mask, rec_maskv;
  int64_t wstream_bytes;
};

// ------------------------------------------------------------------------------------------------
// host: plan + weight packing
// ------------------------------------------------------------------------------------------------
static inline void make_fwd_plan(const MlpLayout& L, FwdPlan& P) {
  P.D = L.D;
  int n = 0, boff = 0;
  auto step_bytes = [](const FwdStep& s) { return (int64_t)s.nslices * s.N * KS * 2 + (int64_t)s.N * 16 * 2; };
  int64_t wbytes = 0;
  for (int i = 0; i < L.D; ++i) {
    FwdStep& s = P.steps[n++];
    s.N = TW; s.epi = EPI_TRUNK; s.bias_src = 0;
    s.last_trunk = (i == L.D - 1);
    s.enc_first = 1;
    if (i == 0) { s.nslices = 2; s.enc_slices = 2; }
    else if (L.skip >= 0 && i == L.skip + 1) { s.nslices = 2 + TW / KS; s.enc_slices = 2; }
    else { s.nslices = TW / KS; s.enc_slices = 0; }
    s.stash_slot = i;
    wbytes += step_bytes(s);
  }
  { FwdStep& s = P.steps[n++];
    s.N = TW; s.epi = EPI_FEATURE; s.bias_src = 1; s.last_trunk = 0;
    s.nslices = TW / KS; s.enc_slices = 0; s.enc_first = 1; s.stash_slot = L.D;
    wbytes += step_bytes(s); }
  { FwdStep& s = P.steps[n++];
    s.N = 128; s.epi = EPI_VIEWS; s.bias_src = 2; s.last_trunk = 0;
    s.nslices = TW / KS + 1; s.enc_slices = 1; s.enc_first = 0; s.stash_slot = -1;
    wbytes += step_bytes(s); }
  P.nsteps = n;
  P.w_alpha_off = boff; boff += TW;
  P.w_rgb_off = boff; boff += 3 * 128;
  P.b_head_off = boff; boff += 4;
  P.bias_off = boff; boff += (L.D + 1) * TW + 128;   // pts_linears[0..D-1], feature_linear, views_linears[0]
  P.nparam_floats = boff;
  P.wstream_bytes = wbytes;
  int off = 0;
  P.rec_encp = off; off += ENC_BYTES;
  P.rec_h0 = off; off += L.D * H_BYTES;
  P.rec_feat = -1;   // the feature layer's output is NOT stashed: the backward never needs it (see mlp_tc_bwd.cu, "T trick")
  P.rec_encv = off; off += ENCV_BYTES;
  P.rec_hv = off; off += HV_BYTES;
  P.rec_mask = off; off += L.D * TROWS * 32;   // 256 relu bits per row per trunk layer: [layer][warpgroup][row][16 B]
  P.rec_maskv = off; off += TROWS * 16;        // 128 bits per row for the views layer: [warpgroup][row][8 B]
  P.rec_bytes = off;
}


// ---- lean MMA issue: descriptors are built once per step; per K=16 MMA only the 14-bit start-address field
// (low word) advances.  hi word = SBO | version, lo word = start>>4 | (LBO>>4)<<16.
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }
__device__ __forceinline__ void umma_f16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                              uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- 1-bit ReLU masks (written by the forward epilogue, applied by the dgrad epilogue) ----------------
// One 32-bit word covers 32 accumulator columns = 16 packed fp16x2 words w = 0..15 (word w = columns 2w, 2w+1).
// Bit = 1: the unit was inactive (negative pre-activation).  The bit of column c sits where ONE byte-permute with
// sign replication turns it into a full fp16 lane mask after a left shift by s = w>>1:
//     position(c) = 15 + 16*(c&1) - 8*(w&1) - s          (w = c>>1)
// i.e. for even w the two bits sit at the sign positions of bytes 1 and 3 of (m << s), for odd w at bytes 0 and 2.
__host__ __device__ constexpr int relu_mask_pos(int c) {
  return 15 + 16 * (c & 1) - 8 * ((c >> 1) & 1) - (c >> 2);
}
// column whose bit lands at position P (inverse of relu_mask_pos)
__host__ __device__ constexpr int relu_mask_col(int P) {
  return 2 * (2 * ((P & 15) < 8 ? 7 - (P & 15) : 15 - (P & 15)) + ((P & 15) < 8 ? 1 : 0)) + (P >> 4);
}
// sign bits of 32 fp32 accumulators in the layout above: 32 funnel shifts, the k-th shifted-in bit ends at 31-k.
// Four independent 8-shift chains (one per result byte) instead of one 32-deep dependency chain: the epilogue warps
// are latency bound, not issue bound.
__device__ __forceinline__ uint32_t shf_in1(uint32_t src, uint32_t acc) {   // (acc << 1) | sign(src)
  uint32_t r;
  asm("shf.l.wrap.b32 %0, %1, %2, 1;" : "=r"(r) : "r"(src), "r"(acc));
  return r;
}
__device__ __forceinline__ uint32_t sign_bits(const uint32_t (&r)[32]) {
  uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    m0 = shf_in1(r[relu_mask_col(31 - j)], m0);
    m1 = shf_in1(r[relu_mask_col(23 - j)], m1);
    m2 = shf_in1(r[relu_mask_col(15 - j)], m2);
    m3 = shf_in1(r[relu_mask_col(7 - j)], m3);
  }
  return (m0 << 24) | (m1 << 16) | (m2 << 8) | m3;
}
// zero the fp16 lanes of packed word w (columns 2w, 2w+1 of the chunk) whose mask bit is set: shift + prmt + lop3
template <int W_>
__device__ __forceinline__ uint32_t relu_mask_apply(uint32_t packed, uint32_t m) {
  uint32_t lanes;
  const uint32_t t = m << (W_ >> 1);
  asm("prmt.b32 %0, %1, %1, %2;" : "=r"(lanes) : "r"(t), "n"((W_ & 1) ? 0xAA88 : 0xBB99));
  return packed & ~lanes;
}
__device__ __forceinline__ bool relu_mask_test(uint32_t m, int c) { return (m >> relu_mask_pos(c)) & 1u; }

// SAT = true: clamp to the finite fp16 range (loss-scaled gradients of the backward chain); SAT = false: overflow to
// inf so that the numerical guard sees it (forward activations, see pack_h2_ovf)
template <bool RELU, bool SAT = true>
__device__ __forceinline__ void pack_chunk(const uint32_t (&r)[32], uint32_t (&pw)[16]) {
#pragma unroll
  for (int w = 0; w < 16; ++w) {
    const float lo = __uint_as_float(r[2 * w]), hi = __uint_as_float(r[2 * w + 1]);
    pw[w] = RELU ? (SAT ? pack_h2_relu(lo, hi) : pack_h2_relu_ovf(lo, hi)) : (SAT ? pack_h2(lo, hi) : pack_h2_ovf(lo, hi));
  }
}
// ---- stash blobs in global memory (activations / gradients kept for the wgrad GEMM) ----------------------
// A [128 x C] fp16 blob is stored as two 64-row halves, each half in the tile-canonical layout with 64 rows:
//     offset(r, c) = (r/64) * (C/8 * 1024) + (c/8) * 1024 + (r%64) * 16 + (c%8) * 2
// so that one 64-sample half of a blob is ONE contiguous run of bytes which is, verbatim, the MN-major UMMA
// operand image the wgrad kernel wants in shared memory (a single bulk copy per operand and stage).
__host__ __device__ constexpr uint32_t stash_off(uint32_t ncg, uint32_t r, uint32_t cg) {
  return (r >> 6) * (ncg * 1024u) + cg * 1024u + (r & 63u) * 16u;
}
// 32 columns (4 groups of 8) of one row of a [128 x 8*ncg] stash blob; c = first column
#ifdef LUSH_EXP_STASH_LOCAL
static __device__ uint8_t g_exp_stash_window[128 * 65536];   // 8 MB: stays in L2
#endif
__device__ __forceinline__ void store_chunk_global(uint8_t* blob, int ncg, int c, int row, const uint32_t (&pw)[16]) {
#ifdef LUSH_EXP_NO_STASH   // timing experiment only (results of the backward are garbage): what do the stash stores cost?
  return;
#endif
#ifdef LUSH_EXP_STASH_LOCAL   // timing experiment only: the same store instructions, but into an L2-resident window per
  blob = g_exp_stash_window + (blockIdx.x & 127) * 65536;   // CTA (no DRAM write traffic): instruction cost vs traffic cost
#endif
#pragma unroll
  for (int g = 0; g < 4; ++g)   // written once, read once by a later kernel: streaming stores (evict-first in L2)
    asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(blob + stash_off(ncg, row, (c >> 3) + g)),
                 "r"(pw[4 * g]), "r"(pw[4 * g + 1]), "r"(pw[4 * g + 2]), "r"(pw[4 * g + 3])
                 : "memory");
}

}  // namespace lush
