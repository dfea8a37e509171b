// Fused positional-encoder + 8x256 scene MLP on the 5th-gen tensor cores (tcgen05 / TMEM), sm_100a.
// impl = 1 of lush_mlp_fwd: fp16 operands (11-bit significand, the same as TF32) with fp32 accumulation.
// Replaces reference helpers:311-344 (Embedder), helpers:394-423 (NeRF.forward) and
// models/lushnerf.py:234-266 (mlpforward) for W = 256 networks.
//
// The kernel is one instance of the chain skeleton in mlp_chain.cuh (read that header first):
//   steps 0..D-1   trunk layers: bias via a ones-column K=16 slice, gamma(x) and the hidden state as shared-memory
//                  A operands; ReLU fused into the fp16 conversion of the epilogue
//   step D         feature_linear (no activation)
//   step D+1       views_linears[0] on [feature | gamma(viewdirs)], N = 128, accumulator region 1
// The sigma head (256->1) and rgb head (128->3) are fp32 register dot products over the accumulators in the
// epilogues; only [M,4] fp32 goes back to HBM -- plus, in training, the fp16 activation stash and 1-bit ReLU
// masks for the backward kernels, written with coalesced 16-byte stores straight from the epilogue registers.
#include "mlp_chain.cuh"

namespace lush {
using namespace tc;

enum FwdEpiKind { FE_TRUNK = 0, FE_TRUNK_LAST = 1, FE_FEATURE = 2, FE_VIEWS = 3 };

// ------------------------------------------------------------------------------------------------
// host: chain plan + weight packing
// ------------------------------------------------------------------------------------------------
static void make_fwd_chain(const MlpLayout& L, const FwdPlan& P, ChainPlan& C) {
  int n = 0;
  for (int i = 0; i < L.D; ++i) {
    ChainStep& s = C.steps[n];
    s.N = TW; s.pre16 = 1; s.enc_src = 0;
    const bool enc = (i == 0) || (L.skip >= 0 && i == L.skip + 1);
    s.enc_slices = enc ? 2 : 0;
    s.h_slices = (i == 0) ? 0 : TW / KS;
    s.wait_a = (i > 0); s.writes_a = 1; s.keep_acc = 0;
    s.free_enc = (L.skip >= 0 ? (i == L.skip + 1) : (i == 0)) ? 1 : 0;   // last step that reads gamma(x)
    s.epi = (i == L.D - 1) ? FE_TRUNK_LAST : FE_TRUNK;
    s.p0 = (int16_t)i; s.p1 = 0;
    s.off0 = P.rec_h0 + i * H_BYTES;
    s.off1 = P.rec_mask + i * TROWS * 32;
    ++n;
  }
  { ChainStep& s = C.steps[n];
    s.N = TW; s.pre16 = 1; s.enc_src = 0; s.enc_slices = 0; s.h_slices = TW / KS;
    s.wait_a = 1; s.writes_a = 1; s.free_enc = 0; s.keep_acc = 0;
    s.epi = FE_FEATURE; s.p0 = 0; s.p1 = 0; s.off0 = P.rec_feat; s.off1 = 0;
    ++n; }
  { ChainStep& s = C.steps[n];
    s.N = 128; s.pre16 = 1; s.enc_src = 1; s.enc_slices = 1; s.h_slices = TW / KS;
    s.wait_a = 1; s.writes_a = 0; s.free_enc = 2; s.keep_acc = 0;
    s.epi = FE_VIEWS; s.p0 = 0; s.p1 = 0; s.off0 = P.rec_hv; s.off1 = P.rec_maskv;
    ++n; }
  C.nsteps = n;
  if (!chain_finish_plan(C, FwdRing::STAGE_BYTES)) C.nsteps = -1;   // rejected by the host entry point
}

// shared-memory map of the forward kernel behind the ring and the activation tile
struct FwdSmem {
  static constexpr int ENC = FwdRing::FREE0;                 // [128 x 64] gamma(x), single buffer
  static constexpr int ENCV = ENC + ENC_BYTES;              // 2 x [128 x 32] gamma(viewdirs)
  static constexpr int ONES = ENCV + 2 * ENCV_BYTES;        // [128 x 16], column 0 = 1 (bias operand)
  static constexpr int PARAMS = ONES + CH_X16_BYTES;        // head weights (fp32)
};

// One block row (blockIdx.y) per step; one thread per packed fp16 element.  The stream of a step holds
//   [bias block N x 16 (column 0 = bias)] [enc block N x 32*enc_slices] [hidden blocks N x 64 ...]
// each block in the tile-canonical layout with N rows: offset(n, k) = (k/8)*(N*16) + n*16 + (k%8)*2.
struct PackJob {
  int64_t out_off;   // byte offset of the step in the stream
  int64_t w_off;     // float offset of the [N, fan_in] weight in the flat params
  int64_t b_off;     // float offset of the bias [N]
  int fan_in, N, pre16, enc_slices, h_slices;
  int enc_real, enc_col0, main_col0;   // nn.Linear column of encoding channel 0 / hidden unit 0
};
struct PackPlan {
  int njobs;
  PackJob jobs[CH_MAX_STEPS];
};

__global__ void pack_fwd_weights_kernel(const float* __restrict__ params, __half* __restrict__ stream,
                                        PackPlan plan) {
  const PackJob& J = plan.jobs[blockIdx.y];
  const int n_pre = J.N * 16 * J.pre16, n_enc = J.N * KS * J.enc_slices;
  const int total = n_pre + n_enc + J.N * KS * J.h_slices;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    int r = e;
    float v = 0.f;
    if (r < n_pre) {
      const int g = r / (J.N * 8), rem = r - g * J.N * 8;
      const int n = rem >> 3, k8 = rem & 7;
      if (g == 0 && k8 == 0) v = params[J.b_off + n];
    } else {
      r -= n_pre;
      int col = -1, n;
      if (r < n_enc) {
        const int kk8 = r / (J.N * 8), rem = r - kk8 * J.N * 8;
        n = rem >> 3;
        const int k = kk8 * 8 + (rem & 7);
        if (k < J.enc_real) col = J.enc_col0 + k;
      } else {
        r -= n_enc;
        const int blk = r / (J.N * 64), rem = r - blk * J.N * 64;
        const int kk8 = rem / (J.N * 8), rem2 = rem - kk8 * J.N * 8;
        n = rem2 >> 3;
        col = J.main_col0 + blk * 64 + kk8 * 8 + (rem2 & 7);
      }
      if (col >= 0) v = params[J.w_off + (int64_t)n * J.fan_in + col];
    }
    stream[J.out_off / 2 + e] = __float2half_rn(v);
  }
}

static void make_pack_plan(const MlpLayout& L, const ChainPlan& C, PackPlan& K) {
  K.njobs = C.nsteps;
  int64_t off = 0;
  for (int s = 0; s < C.nsteps; ++s) {
    const ChainStep& st = C.steps[s];
    PackJob& J = K.jobs[s];
    J.out_off = off;
    J.N = st.N; J.pre16 = st.pre16; J.enc_slices = st.enc_slices; J.h_slices = st.h_slices;
    if (s < L.D) {
      J.w_off = L.pts_w[s]; J.b_off = L.pts_b[s]; J.fan_in = L.pts_in[s];
      J.enc_real = L.in_ch; J.enc_col0 = 0;
      J.main_col0 = (st.enc_slices && s > 0) ? L.in_ch : 0;   // skip layer: [input | hidden] (helpers:401)
    } else if (s == L.D) {
      J.w_off = L.feat_w; J.b_off = L.feat_b; J.fan_in = L.W; J.enc_real = 0; J.enc_col0 = 0; J.main_col0 = 0;
    } else {
      J.w_off = L.views_w; J.b_off = L.views_b; J.fan_in = L.W + L.in_ch_v;   // [feature | views] (helpers:409)
      J.enc_real = L.in_ch_v; J.enc_col0 = L.W; J.main_col0 = 0;
    }
    off += chain_step_bytes(st);
  }
}

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
// gamma(x) for one coordinate triple into the tile-canonical [128 x NCOL] fp16 blob, row `row` (shared memory, and
// the same 16-byte groups to the stash record when gdst != nullptr).
// channel order: [x(3), sin(2^0 x)(3), cos(2^0 x)(3), ...]; columns >= 3+6L are zero padding.
template <int NCOL>
__device__ __forceinline__ void encode_row(uint8_t* blob, uint8_t* gdst, int row, float x0, float x1, float x2,
                                           int L) {
  float vals[NCOL];
#pragma unroll
  for (int c = 0; c < NCOL; ++c) vals[c] = 0.f;
  vals[0] = x0; vals[1] = x1; vals[2] = x2;
  // sin / cos of 2^l x: the exact sincosf at every third octave, the two octaves in between by angle doubling
  // (sin 2a = 2 sin a cos a, cos 2a = 1 - 2 sin^2 a).  At most two doublings from an exact value: <= ~5e-7 absolute,
  // 1/500 of the fp16 rounding (2.4e-4) these values receive two lines below -- and 12 instead of 30 sincosf per point
  // (the exact-fp32 SIMT path, mlp_simt.cu, keeps sincosf at every octave).
  float f = 1.f, s0 = 0.f, c0 = 1.f, s1 = 0.f, c1 = 1.f, s2 = 0.f, c2 = 1.f;
#pragma unroll
  for (int l = 0; l < (NCOL - 3) / 6; ++l) {
    if (l < L) {
      if (l % 3 == 0) {
        sincosf(x0 * f, &s0, &c0);
        sincosf(x1 * f, &s1, &c1);
        sincosf(x2 * f, &s2, &c2);
      } else {
        float t;
        t = 2.f * s0 * c0; c0 = fmaf(-2.f * s0, s0, 1.f); s0 = t;
        t = 2.f * s1 * c1; c1 = fmaf(-2.f * s1, s1, 1.f); s1 = t;
        t = 2.f * s2 * c2; c2 = fmaf(-2.f * s2, s2, 1.f); s2 = t;
      }
      vals[3 + 6 * l + 0] = s0; vals[3 + 6 * l + 1] = s1; vals[3 + 6 * l + 2] = s2;
      vals[3 + 6 * l + 3] = c0; vals[3 + 6 * l + 4] = c1; vals[3 + 6 * l + 5] = c2;
    }
    f *= 2.f;
  }
#pragma unroll
  for (int g = 0; g < NCOL / 8; ++g) {
    uint4 q;
    q.x = pack_h2(vals[8 * g + 0], vals[8 * g + 1]);
    q.y = pack_h2(vals[8 * g + 2], vals[8 * g + 3]);
    q.z = pack_h2(vals[8 * g + 4], vals[8 * g + 5]);
    q.w = pack_h2(vals[8 * g + 6], vals[8 * g + 7]);
    *reinterpret_cast<uint4*>(blob + g * (TROWS * 16) + row * 16) = q;
    if (gdst) *reinterpret_cast<uint4*>(gdst + stash_off(NCOL / 8, row, g)) = q;
  }
}

// ------------------------------------------------------------------------------------------------
// forward kernel
// ------------------------------------------------------------------------------------------------
template <bool TRAIN>
__global__ void __launch_bounds__(CH_THREADS, 1)
mlp_tc_fwd_kernel(const uint8_t* __restrict__ wstream, const float* __restrict__ params_smem_src,
                  const float* __restrict__ ray_batch, int ray_stride, const float* __restrict__ z,
                  float* __restrict__ out, uint8_t* __restrict__ stash, int64_t M, int S, int L_pts,
                  int L_views, ChainPlan C, FwdPlan P) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ ChainBars B;
  __shared__ uint32_t tmem_base_s;
  __shared__ float4 s_part[TROWS];   // warpgroup 0's partial (rgb, sigma) head sums of the current tile

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (M + TROWS - 1) / TROWS;
  float* sparams = reinterpret_cast<float*>(smem + FwdSmem::PARAMS);
  const int nhead = P.bias_off;      // w_alpha | w_rgb | head biases (layer biases ride in the weight stream)

  for (int i = tid; i < nhead; i += CH_THREADS) sparams[i] = params_smem_src[i];
  // ones operand: row r -> [1, 0 x 7 | 0 x 8]: the K=16 slice that adds the bias inside the MMA
  for (int i = tid; i < CH_X16_BYTES / 16; i += CH_THREADS)
    *reinterpret_cast<uint4*>(smem + FwdSmem::ONES + i * 16) = (i < TROWS) ? make_uint4(0x3C00u, 0u, 0u, 0u)
                                                                            : make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) chain_init_bars(B);
  if (warp == CH_WARP_MMA0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == CH_WARP_PRODUCER) {
    if (lane == 0) chain_producer<FwdRing>(B, smem, wstream, C, ntiles);
  } else if (warp >= CH_WARP_MMA0) {
    const ChainMap map = {(uint32_t)FwdSmem::ENC, (uint32_t)FwdSmem::ENCV, (uint32_t)ENCV_BYTES, (uint32_t)FwdSmem::ONES, 0u};
    chain_mma_issuer<FwdRing>(B, smem, tmem_base, C, map, ntiles, (uint32_t)(warp - CH_WARP_MMA0));
  } else if (warp < CH_WG_EPI0) {
    // =========================== prologue warpgroup: sample points + encodings of the NEXT tile ===========
    const int row = tid - CH_WG_PRO * 32;
    int64_t it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int b = (int)(it & 1);
      const int64_t m = tile * TROWS + row;
      float x0 = 0.f, x1 = 0.f, x2 = 0.f, vx = 0.f, vy = 0.f, vz = 0.f;
      if (m < M) {
        const int64_t r = m / S;
        const float* rb = ray_batch + r * ray_stride;
        const float zz = z[m];
        x0 = __fadd_rn(rb[0], __fmul_rn(rb[3], zz));
        x1 = __fadd_rn(rb[1], __fmul_rn(rb[4], zz));
        x2 = __fadd_rn(rb[2], __fmul_rn(rb[5], zz));
        vx = rb[8]; vy = rb[9]; vz = rb[10];
      }
      uint8_t* rec = TRAIN ? stash + tile * (int64_t)P.rec_bytes : nullptr;
      // gamma(viewdirs) is double-buffered; gamma(x) has ONE buffer, free once the previous tile's last trunk
      // layer that reads it (step 0 or the skip layer) has completed -- several layers before that tile ends
      if (it >= 2) mbar_wait(&B.pro_free[b], (uint32_t)(((it >> 1) + 1) & 1));
      encode_row<32>(smem + FwdSmem::ENCV + b * ENCV_BYTES, TRAIN ? rec + P.rec_encv : nullptr, row, vx, vy, vz, L_views);
      if (it >= 1) mbar_wait(&B.enc_free, (uint32_t)((it + 1) & 1));
      encode_row<64>(smem + FwdSmem::ENC, TRAIN ? rec + P.rec_encp : nullptr, row, x0, x1, x2, L_pts);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&B.pro_ready[b]);
    }
  } else if (warp < CH_WARP_PRODUCER) {
    // =========================== epilogue warpgroups ===========================
    // The accumulator is drained one 64-column group per round; warpgroup g converts columns [32g, 32g+32) of the
    // group (one tcgen05.ld.x32, software-pipelined one round ahead): activation, fp16, store into the activation
    // tile (+ stash), publish the group.  The accumulator is released once its last chunk is in registers.
    const int g = (warp - CH_WG_EPI0) >> 2;
    const int row = 32 * (warp & 3) + lane;           // TMEM lane == tile row (a warp may only touch lanes 32*(warp%4)..)
    const uint32_t tlane = (uint32_t)(32 * (warp & 3)) << 16;
    const bool tracer = (warp == CH_WG_EPI0 + 4 * g);
    (void)FwdRing::ACT;   // the activation tile of the ring layout is unused here: activations live in tensor memory
    const float4* wa4 = reinterpret_cast<const float4*>(sparams + P.w_alpha_off);
    const float4* wr4 = reinterpret_cast<const float4*>(sparams + P.w_rgb_off);
    uint32_t d_phase = 0, aux_phase = 0, acc_sel = 0;   // d_phase: bit per accumulator
    int64_t it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int64_t m = tile * TROWS + row;
      uint8_t* rec = TRAIN ? stash + tile * (int64_t)P.rec_bytes : nullptr;
      float alpha = 0.f;
      for (int s = 0; s < C.nsteps; ++s, acc_sel ^= 1u) {   // the forward plan never keeps an accumulator (keep_acc = 0)
        const ChainStep S_ = C.steps[s];
        const int nround = S_.N >> 6;                 // 64-column groups: 256 -> 4, 128 -> 2
        mbar_wait(&B.d_ready[acc_sel], (d_phase >> acc_sel) & 1u); d_phase ^= (1u << acc_sel);
        tc_fence_after();
        if (tracer) chain_trace(C, it, s, 4 + 4 * g);
        const uint32_t tacc = tmem_base + tlane + acc_sel * 256u + 32u * g;
        float c0 = 0.f, c1 = 0.f, c2 = 0.f;
        uint32_t ra[32], rb2[32], pw[16];
        uint32_t mkeep = 0;
        uint2 mpair = make_uint2(0u, 0u);
        tmem_ld32(tacc, ra);
#pragma unroll 1
        for (int rd2 = 0; rd2 < nround; rd2 += 2) {
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {
            const int rd = rd2 + hb;
            uint32_t (&r)[32] = hb ? rb2 : ra;
            const int col = 64 * rd + 32 * g;         // first output column of my chunk
            tmem_ld_wait();
            if (rd + 1 < nround) tmem_ld32(tacc + 64u * (rd + 1), hb ? ra : rb2);
            else {                                    // everything of mine is in registers
              chain_release_acc(B, (int)acc_sel, lane);
              if (tracer) chain_trace(C, it, s, 5 + 4 * g);
            }
            if (S_.epi != FE_VIEWS) {
              // critical path first: fp16 -> activation tile -> publish; everything else rides behind the hand-off
              if (S_.epi == FE_FEATURE) pack_chunk<false, false>(r, pw);
              else pack_chunk<true, false>(r, pw);
              // A operand of the next step: 16 packed columns over the first 16 columns of the very chunk this warp has
              // just loaded -- no other warp ever reads those as accumulator, so the two warpgroups need no barrier
              tmem_st16(tacc + 64u * rd, pw);
              chain_publish_tmem(B, rd, lane);
              if (tracer && (rd >> 1) == g) chain_trace(C, it, s, 6 + 4 * g + (rd & 1));
              if (TRAIN && S_.epi != FE_FEATURE) {   // the feature layer is linear: nothing of it is needed backward
                store_chunk_global(rec + S_.off0, S_.N >> 3, col, row, pw);
                // ReLU bits: one word per round, stored ONCE per step as [warpgroup][row][16 bytes] -- a warp writes 512
                // contiguous bytes (four 4-byte stores per thread at a 32-byte stride cost 7 % of this kernel)
                if (hb == 0) mkeep = sign_bits(r);
                else if (rd2 == 0) { mpair.x = mkeep; mpair.y = sign_bits(r); }
                else *reinterpret_cast<uint4*>(rec + S_.off1 + (g * TROWS + row) * 16) =
                         make_uint4(mpair.x, mpair.y, mkeep, sign_bits(r));
              }
              if (S_.epi == FE_TRUNK_LAST) {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  const float4 w4 = wa4[(col >> 2) + j4];
                  alpha = fmaf(relu_nan(__uint_as_float(r[4 * j4 + 0])), w4.x, alpha);
                  alpha = fmaf(relu_nan(__uint_as_float(r[4 * j4 + 1])), w4.y, alpha);
                  alpha = fmaf(relu_nan(__uint_as_float(r[4 * j4 + 2])), w4.z, alpha);
                  alpha = fmaf(relu_nan(__uint_as_float(r[4 * j4 + 3])), w4.w, alpha);
                }
              }
            } else {
              // views layer: ReLU, rgb head partial sums over my columns; warpgroup 1 combines
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 w0 = wr4[(col >> 2) + j4], w1 = wr4[32 + (col >> 2) + j4], w2 = wr4[64 + (col >> 2) + j4];
                const float v0 = relu_nan(__uint_as_float(r[4 * j4 + 0])), v1 = relu_nan(__uint_as_float(r[4 * j4 + 1]));
                const float v2 = relu_nan(__uint_as_float(r[4 * j4 + 2])), v3 = relu_nan(__uint_as_float(r[4 * j4 + 3]));
                c0 = fmaf(v0, w0.x, c0); c0 = fmaf(v1, w0.y, c0); c0 = fmaf(v2, w0.z, c0); c0 = fmaf(v3, w0.w, c0);
                c1 = fmaf(v0, w1.x, c1); c1 = fmaf(v1, w1.y, c1); c1 = fmaf(v2, w1.z, c1); c1 = fmaf(v3, w1.w, c1);
                c2 = fmaf(v0, w2.x, c2); c2 = fmaf(v1, w2.y, c2); c2 = fmaf(v2, w2.z, c2); c2 = fmaf(v3, w2.w, c2);
              }
              if (TRAIN) {
                pack_chunk<true, false>(r, pw);
                store_chunk_global(rec + S_.off0, S_.N >> 3, col, row, pw);
                // ReLU bits of the 128-wide views layer: [warpgroup][row][8 bytes]
                if (hb == 0) mkeep = sign_bits(r);
                else *reinterpret_cast<uint2*>(rec + S_.off1 + (g * TROWS + row) * 8) = make_uint2(mkeep, sign_bits(r));
              }
            }
          }
        }
        if (S_.epi == FE_VIEWS) {
          if (g == 0) {
            s_part[row] = make_float4(c0, c1, c2, alpha);
            __syncwarp();
            if (lane == 0) mbar_arrive(&B.aux);
          } else {
            mbar_wait(&B.aux, aux_phase); aux_phase ^= 1;
            const float4 o = s_part[row];
            if (m < M) {
              const float* bh = sparams + P.b_head_off;
              reinterpret_cast<float4*>(out)[m] =
                  make_float4(c0 + o.x + bh[1], c1 + o.y + bh[2], c2 + o.z + bh[3], alpha + o.w + bh[0]);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == CH_WARP_MMA0) tmem_dealloc(tmem_base, 512);
}

static size_t fwd_smem_bytes(const FwdPlan& P) { return (size_t)FwdSmem::PARAMS + (size_t)P.bias_off * 4; }

// head weights gathered into one contiguous float block: w_alpha[256] | w_rgb[3][128] | b_alpha, b_rgb[3]
__global__ void gather_fwd_params_kernel(const float* __restrict__ params, float* __restrict__ dst, MlpLayout L,
                                         FwdPlan P) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.nparam_floats) return;
  float v;
  if (i < P.w_rgb_off) v = params[L.alpha_w + (i - P.w_alpha_off)];
  else if (i < P.b_head_off) v = params[L.rgb_w + (i - P.w_rgb_off)];
  else if (i < P.bias_off) {
    int c = i - P.b_head_off;
    v = (c == 0) ? params[L.alpha_b] : params[L.rgb_b + c - 1];
  } else {
    const int c = i - P.bias_off, s = c / TW, n = c - s * TW;
    v = (s < L.D) ? params[L.pts_b[s] + n] : (s == L.D ? params[L.feat_b + n] : params[L.views_b + n]);
  }
  dst[i] = v;
}

}  // namespace lush

using namespace lush;

size_t lush_mlp_tc_bwd_workspace_bytes(const MlpLayout& L, int64_t R, int S);   // mlp_tc_bwd.cu

bool lush_mlp_tc_training_supported() { return true; }

bool lush_mlp_tc_supported(const MlpLayout& L) {
  return L.W == TW && L.out_mode == 0 && L.D >= 2 && L.D <= kMaxDepth && L.in_ch <= 64 && L.in_ch_v <= 32 &&
         (L.skip < 0 || (L.skip >= 0 && L.skip < L.D - 1));
}

size_t lush_mlp_tc_stash_bytes(const MlpLayout& L, int64_t R, int S) {
  FwdPlan P;
  make_fwd_plan(L, P);
  const int64_t ntiles = (R * S + TROWS - 1) / TROWS;
  return (size_t)ntiles * P.rec_bytes;
}

// forward workspace: [packed weight stream | gathered params]; backward: see mlp_tc_bwd.cu
size_t lush_mlp_tc_workspace_bytes(const MlpLayout& L, int64_t R, int S, int training) {
  if (training) return lush_mlp_tc_bwd_workspace_bytes(L, R, S);
  FwdPlan P;
  make_fwd_plan(L, P);
  ChainPlan C;
  make_fwd_chain(L, P, C);
  return (size_t)((C.stream_bytes + 255) / 256) * 256 + (size_t)P.nparam_floats * 4 + 256;
}

// biases + head weights gathered into one float block (also used by the backward kernels)
int lush_tc_gather_head_params(const MlpLayout& L, const float* params, float* dst, cudaStream_t st) {
  FwdPlan P;
  make_fwd_plan(L, P);
  gather_fwd_params_kernel<<<(P.nparam_floats + 255) / 256, 256, 0, st>>>(params, dst, L, P);
  LUSH_LAUNCH_CHECK("gather_fwd_params");
  return 0;
}

int lush_mlp_tc_fwd(const MlpLayout& L, const lush_mlp_desc* desc, const float* params, const float* ray_batch,
                    int ray_stride, const float* z, float* out, void* stash, void* workspace,
                    size_t workspace_bytes, int64_t R, int S, cudaStream_t st) {
  FwdPlan P;
  make_fwd_plan(L, P);
  const size_t need = lush_mlp_tc_workspace_bytes(L, R, S, 0);
  if (!workspace || workspace_bytes < need) {
    set_error("mlp_fwd(tc): workspace too small (%zu < %zu)", workspace_bytes, need);
    return LUSH_EWORKSPACE;
  }
  LUSH_REQUIRE(((uintptr_t)workspace & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)stash & 15) == 0,
               "mlp_fwd(tc): unaligned buffers");
  ChainPlan C;
  make_fwd_chain(L, P, C);
  LUSH_REQUIRE(C.nsteps > 0, "mlp_fwd(tc): a weight block of this topology does not fit the ring stage");
  __half* wstream = reinterpret_cast<__half*>(workspace);
  float* gparams = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + ((C.stream_bytes + 255) / 256) * 256);
  PackPlan K;
  make_pack_plan(L, C, K);
  pack_fwd_weights_kernel<<<dim3(64, K.njobs), 256, 0, st>>>(params, wstream, K);
  LUSH_LAUNCH_CHECK("pack_fwd_weights");
  int rc = lush_tc_gather_head_params(L, params, gparams, st);
  if (rc) return rc;
  const int64_t M = R * S;
  const int64_t ntiles = (M + TROWS - 1) / TROWS;
  const size_t smem = fwd_smem_bytes(P);
  LUSH_REQUIRE(smem + 1024 <= 227 * 1024, "mlp_fwd(tc): shared memory budget exceeded (%zu)", smem);
  int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
  if (stash != nullptr) {
    rc = check_cuda(cudaFuncSetAttribute(mlp_tc_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem), "mlp_fwd(tc): smem attr");
    if (rc) return rc;
    mlp_tc_fwd_kernel<true><<<grid, CH_THREADS, smem, st>>>((const uint8_t*)wstream, gparams, ray_batch, ray_stride,
                                                           z, out, (uint8_t*)stash, M, S, desc->L_pts,
                                                           desc->L_views, C, P);
  } else {
    rc = check_cuda(cudaFuncSetAttribute(mlp_tc_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem), "mlp_fwd(tc): smem attr");
    if (rc) return rc;
    mlp_tc_fwd_kernel<false><<<grid, CH_THREADS, smem, st>>>((const uint8_t*)wstream, gparams, ray_batch, ray_stride,
                                                            z, out, nullptr, M, S, desc->L_pts, desc->L_views, C,
                                                            P);
  }
  LUSH_LAUNCH_CHECK("mlp_tc_fwd");
  return 0;
}

extern "C" int lush_debug_read_chain_trace(long long* host_out, int n) {   // forward kernel's trace (this translation unit)
  using namespace lush;
  LUSH_REQUIRE(host_out && n > 0 && n <= CH_MAX_STEPS * 16, "read_chain_trace: bad arguments");
  return check_cuda(cudaMemcpyFromSymbol(host_out, g_chain_trace, sizeof(long long) * n), "read_chain_trace");
}
