// Fused positional-encoder + 8x256 scene MLP on the 5th-gen tensor cores (tcgen05 / TMEM), sm_100a.
// impl = 1 of lush_mlp_fwd: fp16 operands (11-bit significand, the same as TF32) with fp32 accumulation.
// Replaces reference helpers:311-344 (Embedder), helpers:394-423 (NeRF.forward) and
// models/lushnerf.py:234-266 (mlpforward) for W = 256 networks.
//
// Forward kernel (persistent, one CTA per SM, 320 threads):
//   warp 0      weight producer: streams the pre-packed fp16 weight K-slices (16 KB each) with 1-D bulk
//               async copies (cp.async.bulk, the TMA engine) through a 3-stage mbarrier ring
//   warp 1      MMA issuer: one thread issues tcgen05.mma (M=128, N=256|128, K=16) for TWO 128-row sample
//               tiles per weight slice (halves L2 weight traffic per row), accumulators in TMEM (2 x 256 cols)
//   warps 2-9   two epilogue warpgroups (one per tile): tcgen05.ld -> +bias -> ReLU -> fp16 -> st.shared in
//               the UMMA operand layout, i.e. the next layer's A operand never leaves the SM.  The sigma
//               head (256->1) and rgb head (128->3) are register dot products in the epilogues; only
//               [M,4] fp32 goes back to HBM (plus, in training, the fp16 activation stash for backward).
// Activations: H tile [128 x 256] fp16 (64 KB) per tile, encodings [128 x 64] (16 KB) per tile.
#include "mlp_tc.cuh"

namespace lush {
using namespace tc;

// One thread per packed fp16 element of the forward weight stream.
// Step s = nslices regular slices [N x 32] followed by one bias slice [N x 16] (column 0 = bias).
// element (n, kk) of slice j: offset = step_base + j*N*KS*2 + (kk/8)*(N*16) + n*16 + (kk%8)*2
struct PackJob {
  int64_t out_off;   // byte offset of the step in the stream
  int64_t w_off;     // float offset of the [N, fan_in] weight in the flat params
  int64_t b_off;     // float offset of the bias [N]
  int N, fan_in, nslices;
  int enc_slices, enc_first, enc_real, enc_col0, main_col0;  // column mapping (see pack kernel)
};
struct PackPlan {
  int njobs;
  PackJob jobs[MAX_STEPS];
};

__global__ void pack_fwd_weights_kernel(const float* __restrict__ params, __half* __restrict__ stream,
                                        PackPlan plan) {
  const PackJob& J = plan.jobs[blockIdx.y];
  const int64_t nreg = (int64_t)J.nslices * J.N * KS;
  const int64_t total = nreg + (int64_t)J.N * 16;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    float v = 0.f;
    if (e < nreg) {
      int j = (int)(e / (J.N * KS));
      int rem = (int)(e - (int64_t)j * J.N * KS);
      int kk8 = rem / (J.N * 8);          // which 8-column group inside the slice
      int rem2 = rem - kk8 * J.N * 8;
      int n = rem2 / 8, k8 = rem2 % 8;
      int k = j * KS + kk8 * 8 + k8;      // K index within the step (padded, concatenated segments)
      int col = -1;                       // nn.Linear column
      const int enc_k = J.enc_slices * KS;
      const int main_k = (J.nslices - J.enc_slices) * KS;
      if (J.enc_first) {
        if (k < enc_k) { if (k < J.enc_real) col = J.enc_col0 + k; }
        else col = J.main_col0 + (k - enc_k);
      } else {
        if (k < main_k) col = J.main_col0 + k;
        else if (k - main_k < J.enc_real) col = J.enc_col0 + (k - main_k);
      }
      if (col >= 0) v = params[J.w_off + (int64_t)n * J.fan_in + col];
    } else {
      int rem = (int)(e - nreg);          // bias slice: [N x 16], only column 0 is non-zero
      int g = rem / (J.N * 8);
      int rem2 = rem - g * J.N * 8;
      int n = rem2 / 8, k8 = rem2 % 8;
      if (g == 0 && k8 == 0) v = params[J.b_off + n];
    }
    stream[J.out_off / 2 + e] = __float2half_rn(v);
  }
}

static void make_pack_plan(const MlpLayout& L, const FwdPlan& P, PackPlan& K) {
  K.njobs = P.nsteps;
  int64_t off = 0;
  for (int s = 0; s < P.nsteps; ++s) {
    const FwdStep& st = P.steps[s];
    PackJob& J = K.jobs[s];
    J.out_off = off;
    J.N = st.N; J.nslices = st.nslices; J.enc_slices = st.enc_slices; J.enc_first = st.enc_first;
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
    off += (int64_t)st.nslices * st.N * KS * 2 + (int64_t)st.N * 16 * 2;
  }
}

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
struct FwdSmem {
  // offsets into dynamic shared memory
  static constexpr int H0 = 0;
  static constexpr int ENC0 = 2 * H_BYTES;
  static constexpr int STAGE0 = ENC0 + 2 * ENC_BYTES;
  static constexpr int ONES = STAGE0 + NSTAGE * STAGE_BYTES;   // [128 x 16] fp16, column 0 = 1 (bias operand)
  static constexpr int PARAMS = ONES + TROWS * 16 * 2;
};

// gamma(x) for one coordinate triple into the tile-canonical [128 x 64] (or [128 x 32]) fp16 blob, row `row`.
// channel order: [x(3), sin(2^0 x)(3), cos(2^0 x)(3), ...]; columns >= 3+6L are zero padding.
template <int NCOL>
__device__ __forceinline__ void encode_row_to_smem(uint8_t* blob, int row, float x0, float x1, float x2, int L) {
  float vals[NCOL];
#pragma unroll
  for (int c = 0; c < NCOL; ++c) vals[c] = 0.f;
  vals[0] = x0; vals[1] = x1; vals[2] = x2;
  float f = 1.f;
#pragma unroll
  for (int l = 0; l < (NCOL - 3) / 6; ++l) {
    if (l < L) {
      float s0, c0, s1, c1, s2, c2;
      sincosf(x0 * f, &s0, &c0);
      sincosf(x1 * f, &s1, &c1);
      sincosf(x2 * f, &s2, &c2);
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
  }
}

// 32 accumulator columns -> fp16 (optionally through ReLU) -> 4 x 16-byte stores in the operand layout
template <bool RELU>
__device__ __forceinline__ void store_chunk(uint8_t* Hb, int cc, int row, const uint32_t (&r)[32]) {
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint4 q;
    if (RELU) {
      q.x = pack_h2_relu(__uint_as_float(r[8 * g + 0]), __uint_as_float(r[8 * g + 1]));
      q.y = pack_h2_relu(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3]));
      q.z = pack_h2_relu(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5]));
      q.w = pack_h2_relu(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7]));
    } else {
      q.x = pack_h2(__uint_as_float(r[8 * g + 0]), __uint_as_float(r[8 * g + 1]));
      q.y = pack_h2(__uint_as_float(r[8 * g + 2]), __uint_as_float(r[8 * g + 3]));
      q.z = pack_h2(__uint_as_float(r[8 * g + 4]), __uint_as_float(r[8 * g + 5]));
      q.w = pack_h2(__uint_as_float(r[8 * g + 6]), __uint_as_float(r[8 * g + 7]));
    }
    *reinterpret_cast<uint4*>(Hb + ((cc >> 3) + g) * (TROWS * 16) + row * 16) = q;
  }
}
// ------------------------------------------------------------------------------------------------
// forward kernel
// ------------------------------------------------------------------------------------------------
template <bool TRAIN>
__global__ void __launch_bounds__(FWD_THREADS, 1)
mlp_tc_fwd_kernel(const __half* __restrict__ wstream, const float* __restrict__ params_smem_src,
                  const float* __restrict__ ray_batch, int ray_stride, const float* __restrict__ z,
                  float* __restrict__ out, uint8_t* __restrict__ stash, int64_t M, int S, int L_pts,
                  int L_views, FwdPlan P) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full_bar[NSTAGE], empty_bar[NSTAGE], a_ready[2], d_ready[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (M + TROWS - 1) / TROWS;
  const int64_t nsuper = (ntiles + 1) / 2;
  float* sparams = reinterpret_cast<float*>(smem + FwdSmem::PARAMS);

  for (int i = tid; i < P.nparam_floats; i += FWD_THREADS) sparams[i] = params_smem_src[i];
  for (int i = tid; i < TROWS * 16 * 2 / 16; i += FWD_THREADS)   // ones operand: row r -> [1,0,...,0 | 0 x 8]
    *reinterpret_cast<uint4*>(smem + FwdSmem::ONES + i * 16) = (i < TROWS) ? make_uint4(0x3C00u, 0u, 0u, 0u)
                                                                            : make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 2); }   // both issuers release a stage
    for (int t = 0; t < 2; ++t) { mbar_init(&a_ready[t], 4 * 32); mbar_init(&d_ready[t], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // =========================== weight producer ===========================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int64_t st = blockIdx.x; st < nsuper; st += gridDim.x) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(wstream);
        for (int s = 0; s < P.nsteps; ++s) {
          const int ns = P.steps[s].nslices;
          for (int j = 0; j <= ns; ++j) {   // j == ns: the bias slice (K = 16)
            const uint32_t bytes = (uint32_t)P.steps[s].N * (j < ns ? KS : 16) * 2;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], bytes);
            bulk_g2s(smem + FwdSmem::STAGE0 + stage * STAGE_BYTES, src, bytes, &full_bar[stage]);
            src += bytes;
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // =========================== MMA issuers: warp 1 -> tile slot 0, warp 2 -> tile slot 1 =================
    if (lane == 0) {
      const int t = warp - 1;
      int stage = 0; uint32_t phase = 0, a_phase = 0;
      const uint32_t sbase = smem_u32(smem);
      const uint32_t tacc = tmem_base + t * 256;
      const uint32_t a_hi = desc_hi(128u), b_hi = desc_hi(128u);
      const uint32_t h_lo = desc_lo(sbase + FwdSmem::H0 + t * H_BYTES, TROWS * 16u);
      const uint32_t e_lo = desc_lo(sbase + FwdSmem::ENC0 + t * ENC_BYTES, TROWS * 16u);
      const uint32_t ones_lo = desc_lo(sbase + FwdSmem::ONES, TROWS * 16u);
      constexpr uint32_t A_K16 = (2u * TROWS * 16u) >> 4;       // start-address units per 16 columns of A
      for (int64_t st = blockIdx.x; st < nsuper; st += gridDim.x) {
        const bool active = (2 * st + t) < ntiles;
        for (int s = 0; s < P.nsteps; ++s) {
          const int N = P.steps[s].N, ns = P.steps[s].nslices, es = P.steps[s].enc_slices;
          const int first_enc = P.steps[s].enc_first ? 0 : ns - es;   // slices [first_enc, first_enc+es) read ENC
          const uint32_t idesc = make_idesc_f16(TROWS, N, 0, 0);
          const uint32_t b_lbo = (uint32_t)N * 16u;
          const uint32_t B_K16 = (2u * b_lbo) >> 4;
          int hcol = 0, ecol = 0;    // running 16-column offsets inside H / ENC
          for (int j = 0; j <= ns; ++j) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (active) {
              const uint32_t b_lo = desc_lo(sbase + FwdSmem::STAGE0 + stage * STAGE_BYTES, b_lbo);
              if (j == ns) {
                umma_f16_lohi(tacc, ones_lo, a_hi, b_lo, b_hi, idesc, 1u);   // + bias
              } else {
                if (j == 0) { mbar_wait(&a_ready[t], a_phase); a_phase ^= 1; tc_fence_after(); }
                const bool from_enc = (j >= first_enc) && (j < first_enc + es);
                uint32_t a_lo;
                if (from_enc) { a_lo = e_lo + ecol * A_K16; ecol += 2; } else { a_lo = h_lo + hcol * A_K16; hcol += 2; }
                umma_f16_lohi(tacc, a_lo, a_hi, b_lo, b_hi, idesc, j > 0 ? 1u : 0u);
                umma_f16_lohi(tacc, a_lo + A_K16, a_hi, b_lo + B_K16, b_hi, idesc, 1u);
              }
              umma_commit(&empty_bar[stage]);
            } else {
              mbar_arrive(&empty_bar[stage]);
            }
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
          if (active) umma_commit(&d_ready[t]);
        }
      }
    }
  } else {
    // =========================== epilogue warpgroups ===========================
    const int t = (warp - EPI_WARP0) >> 2;            // tile slot handled by this warpgroup
    const int row = 32 * (warp & 3) + lane;           // TMEM lane == tile row (a warp may only touch lanes 32*(warp%4)..)
    const uint32_t tlane = (uint32_t)(32 * (warp & 3)) << 16;
    uint8_t* Hb = smem + FwdSmem::H0 + t * H_BYTES;
    uint8_t* Eb = smem + FwdSmem::ENC0 + t * ENC_BYTES;
    const bool leader = ((warp - EPI_WARP0) & 3) == 0 && lane == 0;
    const int bar_id = 1 + t;
    const float4* wa4 = reinterpret_cast<const float4*>(sparams + P.w_alpha_off);
    const float* wr = sparams + P.w_rgb_off;
    uint32_t d_phase = 0;
    for (int64_t st = blockIdx.x; st < nsuper; st += gridDim.x) {
      const int64_t tile = 2 * st + t;
      if (tile >= ntiles) continue;
      const int64_t m = tile * TROWS + row;
      const bool valid = m < M;
      uint8_t* rec = TRAIN ? stash + tile * (int64_t)P.rec_bytes : nullptr;
      const uint32_t tacc = tmem_base + tlane + t * 256;
      // ---- prologue: sample point + encodings
      float vx = 0.f, vy = 0.f, vz = 0.f;
      {
        float x0 = 0.f, x1 = 0.f, x2 = 0.f;
        if (valid) {
          const int64_t r = m / S;
          const float* rb = ray_batch + r * ray_stride;
          const float zz = z[m];
          x0 = __fadd_rn(rb[0], __fmul_rn(rb[3], zz));
          x1 = __fadd_rn(rb[1], __fmul_rn(rb[4], zz));
          x2 = __fadd_rn(rb[2], __fmul_rn(rb[5], zz));
          vx = rb[8]; vy = rb[9]; vz = rb[10];
        }
        if (TRAIN) {                                   // previous tile's stash stores still read H/ENC
          if (leader) bulk_wait_read_all();
          named_bar_sync(bar_id, 128);
        }
        encode_row_to_smem<64>(Eb, row, x0, x1, x2, L_pts);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&a_ready[t]);
      if (TRAIN) {
        named_bar_sync(bar_id, 128);
        if (leader) { bulk_s2g(rec + P.rec_encp, Eb, ENC_BYTES); bulk_commit(); }
      }
      float alpha = 0.f;
      for (int s = 0; s < P.nsteps; ++s) {
        const FwdStep& S_ = P.steps[s];
        mbar_wait(&d_ready[t], d_phase); d_phase ^= 1;
        tc_fence_after();
        uint32_t ra[32], rb2[32];
        tmem_ld32(tacc, ra);                           // overlaps with the stash bookkeeping below
        if (TRAIN) {                                   // the H blob of the previous step is being stored
          if (leader) bulk_wait_read_all();
          named_bar_sync(bar_id, 128);
        }
        if (S_.epi == EPI_VIEWS) {
          float c0 = 0.f, c1 = 0.f, c2 = 0.f;
          uint32_t mbits[4];
#pragma unroll
          for (int cc = 0; cc < 128; cc += 64) {
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
              uint32_t (&r)[32] = hb ? rb2 : ra;
              const int c = cc + 32 * hb;
              tmem_ld_wait();
              if (c + 32 < 128) tmem_ld32(tacc + c + 32, hb ? ra : rb2);
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float v = fmaxf(__uint_as_float(r[j]), 0.f);
                c0 = fmaf(v, wr[c + j], c0);
                c1 = fmaf(v, wr[128 + c + j], c1);
                c2 = fmaf(v, wr[256 + c + j], c2);
              }
              if (TRAIN) {
                mbits[c >> 5] = sign_bits(r);
                store_chunk<true>(Hb, c, row, r);
              }
            }
          }
          if (valid) {
            const float* bh = sparams + P.b_head_off;
            float4 o = make_float4(c0 + bh[1], c1 + bh[2], c2 + bh[3], alpha + bh[0]);
            reinterpret_cast<float4*>(out)[m] = o;
          }
          if (TRAIN) {
            *reinterpret_cast<uint4*>(rec + P.rec_maskv + row * 16) = make_uint4(mbits[0], mbits[1], mbits[2], mbits[3]);
            fence_proxy_async();
            named_bar_sync(bar_id, 128);
            if (leader) { bulk_s2g(rec + P.rec_hv, Hb, HV_BYTES); bulk_commit(); }
          }
          tc_fence_before();   // orders our TMEM reads before the next tile's a_ready arrive
        } else {
          uint32_t mbits[8];
#pragma unroll
          for (int cc = 0; cc < TW; cc += 64) {
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
              uint32_t (&r)[32] = hb ? rb2 : ra;
              const int c = cc + 32 * hb;
              tmem_ld_wait();
              if (c + 32 < TW) tmem_ld32(tacc + c + 32, hb ? ra : rb2);
              if (TRAIN && S_.epi == EPI_TRUNK) mbits[c >> 5] = sign_bits(r);
              if (S_.last_trunk) {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  const float4 w4 = wa4[(c >> 2) + j4];
                  alpha = fmaf(fmaxf(__uint_as_float(r[4 * j4 + 0]), 0.f), w4.x, alpha);
                  alpha = fmaf(fmaxf(__uint_as_float(r[4 * j4 + 1]), 0.f), w4.y, alpha);
                  alpha = fmaf(fmaxf(__uint_as_float(r[4 * j4 + 2]), 0.f), w4.z, alpha);
                  alpha = fmaf(fmaxf(__uint_as_float(r[4 * j4 + 3]), 0.f), w4.w, alpha);
                }
              }
              if (S_.epi == EPI_TRUNK) store_chunk<true>(Hb, c, row, r);
              else store_chunk<false>(Hb, c, row, r);
            }
          }
          if (S_.epi == EPI_FEATURE)   // the encoding buffer is dead after the trunk: reuse it for gamma(viewdirs)
            encode_row_to_smem<32>(Eb, row, vx, vy, vz, L_views);
          if (TRAIN && S_.epi == EPI_TRUNK) {
            uint4* mp = reinterpret_cast<uint4*>(rec + P.rec_mask + (s * TROWS + row) * 32);
            mp[0] = make_uint4(mbits[0], mbits[1], mbits[2], mbits[3]);
            mp[1] = make_uint4(mbits[4], mbits[5], mbits[6], mbits[7]);
          }
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(&a_ready[t]);
          if (TRAIN) {
            named_bar_sync(bar_id, 128);
            if (leader) {
              bulk_s2g(rec + P.rec_h0 + S_.stash_slot * H_BYTES, Hb, H_BYTES);
              if (S_.epi == EPI_FEATURE) bulk_s2g(rec + P.rec_encv, Eb, ENCV_BYTES);
              bulk_commit();
            }
          }
        }
      }
    }
    if (TRAIN && leader) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static size_t fwd_smem_bytes(const FwdPlan& P) { return (size_t)FwdSmem::PARAMS + (size_t)P.nparam_floats * 4; }

// head weights gathered into one contiguous float block: w_alpha[256] | w_rgb[3][128] | b_alpha, b_rgb[3]
__global__ void gather_fwd_params_kernel(const float* __restrict__ params, float* __restrict__ dst, MlpLayout L,
                                         FwdPlan P) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.nparam_floats) return;
  float v;
  if (i < P.w_rgb_off) v = params[L.alpha_w + (i - P.w_alpha_off)];
  else if (i < P.b_head_off) v = params[L.rgb_w + (i - P.w_rgb_off)];
  else {
    int c = i - P.b_head_off;
    v = (c == 0) ? params[L.alpha_b] : params[L.rgb_b + c - 1];
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
  return (size_t)((P.wstream_bytes + 255) / 256) * 256 + (size_t)P.nparam_floats * 4 + 256;
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
  __half* wstream = reinterpret_cast<__half*>(workspace);
  float* gparams = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + ((P.wstream_bytes + 255) / 256) * 256);
  PackPlan K;
  make_pack_plan(L, P, K);
  pack_fwd_weights_kernel<<<dim3(64, K.njobs), 256, 0, st>>>(params, wstream, K);
  LUSH_LAUNCH_CHECK("pack_fwd_weights");
  int rc = lush_tc_gather_head_params(L, params, gparams, st);
  if (rc) return rc;
  const int64_t M = R * S;
  const int64_t ntiles = (M + TROWS - 1) / TROWS, nsuper = (ntiles + 1) / 2;
  const size_t smem = fwd_smem_bytes(P);
  LUSH_REQUIRE(smem + 256 <= 227 * 1024, "mlp_fwd(tc): shared memory budget exceeded (%zu)", smem);
  int grid = (int)(nsuper < kNumSMs ? nsuper : kNumSMs);
  if (stash != nullptr) {
    rc = check_cuda(cudaFuncSetAttribute(mlp_tc_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem), "mlp_fwd(tc): smem attr");
    if (rc) return rc;
    mlp_tc_fwd_kernel<true><<<grid, FWD_THREADS, smem, st>>>(wstream, gparams, ray_batch, ray_stride, z, out,
                                                            (uint8_t*)stash, M, S, desc->L_pts, desc->L_views, P);
  } else {
    rc = check_cuda(cudaFuncSetAttribute(mlp_tc_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem), "mlp_fwd(tc): smem attr");
    if (rc) return rc;
    mlp_tc_fwd_kernel<false><<<grid, FWD_THREADS, smem, st>>>(wstream, gparams, ray_batch, ray_stride, z, out,
                                                             nullptr, M, S, desc->L_pts, desc->L_views, P);
  }
  LUSH_LAUNCH_CHECK("mlp_tc_fwd");
  return 0;
}
