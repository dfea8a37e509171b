// Backward of the fused scene MLP on tcgen05 (impl = 1 of lush_mlp_bwd), sm_100a.
// Replaces the autograd graph of reference helpers:394-423 (mm x2 per Linear, threshold_backward, bias sums).
//
// Three stages over the fp16 activation stash written by the forward kernel (mlp_tc.cu):
//  1. dgrad  (mlp_tc_dgrad_kernel): per 128-sample tile the chain dZ_L -> ... -> dZ_0 stays on chip exactly
//     like the forward: weights (W^T, fp16, pre-packed) stream through a 3-stage bulk-copy ring, one thread
//     issues tcgen05.mma for two tiles per slice, epilogue warpgroups apply the 1-bit ReLU masks and write
//     the next A operand to shared memory; every dZ tile is also bulk-stored to HBM for stage 2.
//     Gradients are carried in fp16 under a dynamic power-of-two loss scale (max|g_out| -> 32).
//  2. wgrad  (mlp_tc_wgrad_kernel): dW_l = dZ_l^T . X_l with the reduction over samples; both operands are
//     the SAME stash blobs read as MN-major UMMA operands; fp32 accumulators stay in TMEM across a chunk of
//     tiles and are flushed once per chunk with fp32 atomics into the flat gradient buffer.  Bias gradients
//     are column sums taken from the dZ stage in shared memory by the otherwise idle epilogue warps.
//  3. small SIMT kernels: head gradients (sigma / rgb heads), encoder backward, per-ray reduction of d/d(o, d, viewdirs).
#include <algorithm>
#include "mlp_chain.cuh"

namespace lush {

constexpr float kGradTarget = 32.f;   // scaled max |g_out|

enum BwdEpi { BE_VIEWS = 0, BE_FEAT = 1, BE_TRUNK = 2, BE_ENC = 3 };

// gstash record (per 128-sample tile): dZv [128 x 128] | dZ_0 .. dZ_{D-1} [128 x 256] each
//
// The "T trick".  The feature layer is linear (feat = H_{D-1} Wf^T + bf feeds the views layer directly), so with
//   T = dZv^T H_{D-1}  [128 x 256]   and   sv = sum_s dZv  [128]        (Wv' = the feature columns of the views weight)
//   dW_f       = Wv'^T T              db_f = Wv'^T sv
//   dW_v[:, feature columns] = T Wf^T + sv bf^T
// ONE wgrad job (dZv, H_{D-1}) replaces the two jobs (dZf, H_{D-1}) and (dZv, feat): neither feat (forward) nor
// dZf = dZv Wv' (dgrad) is written to HBM or read back -- 256 KB less traffic per 128-sample tile (13 %) -- and two
// 128 x 256 x 256 products per step (fv_post_kernel) finish the job.  dZf still exists on chip as a link of the chain.
struct BwdPlan {
  int grec_bytes, grec_v, grec_l0;
  int has_enc;          // d/d gamma(x) is produced (two [128 x 64] fp16 blobs per tile: skip layer, layer 0)
  int has_skip_enc;     // ... and the skip-layer blob (slot 0) exists
};

struct BPackJob {
  int64_t out_off;      // byte offset of the step in the stream
  int64_t w_off;        // float offset of the [rows, fan_in] weight whose TRANSPOSE is the B operand
  int fan_in, col0, ncols_real, N, K;   // B[n][k] = W[k][col0 + n] for n < ncols_real, k < K
  int pre_kind;         // 0: none, 1: rgb head rows (k < 3 -> w_rgb[k][n]), 2: sigma head (k == 3 -> w_alpha[n])
  int64_t pre_off;      // float offset of that head weight
};
struct BPackPlan {
  int njobs;
  BPackJob jobs[CH_MAX_STEPS];
};

// dgrad chain:  P: dZv = mask_v(G16 . Wrgb)            N=128  K=16 (pre16 only)
//               0: g_feat = dZv . Wv[:, :W]             N=256  K=128
//               1: dZ_{D-1} = mask(dZf . Wf + g_sigma w_alpha)     N=256  K=256 + pre16
//               l = D-1..1: [skip layer: d gamma += dZ_l . W_l[:, :in_ch]  N=64]  dZ_{l-1} = mask(dZ_l . W_l[:, hidden])
//               last: d gamma += dZ_0 . W_0            N=64
static void make_bwd_chain(const MlpLayout& L, bool need_in, ChainPlan& C, BwdPlan& B, BPackPlan& K) {
  B.grec_v = 0;
  B.grec_l0 = HV_BYTES;
  B.grec_bytes = B.grec_l0 + L.D * H_BYTES;
  B.has_enc = need_in ? 1 : 0;
  B.has_skip_enc = 0;
  int n = 0;
  int64_t off = 0;
  auto add = [&](int N, int pre16, int h_slices, int wait_a, int writes_a, int free_enc, int epi, int p0, int off0,
                 int64_t w_off, int fan_in, int col0, int ncols_real, int pre_kind, int64_t pre_off) {
    ChainStep& s = C.steps[n];
    s.N = (int16_t)N; s.pre16 = (int8_t)pre16; s.enc_slices = 0; s.enc_src = 0; s.h_slices = (int8_t)h_slices;
    s.wait_a = (int8_t)wait_a; s.writes_a = (int8_t)writes_a; s.free_enc = (int8_t)free_enc; s.epi = (int8_t)epi;
    s.keep_acc = 0;
    s.p0 = (int16_t)p0; s.p1 = 0; s.off0 = off0; s.off1 = 0;
    BPackJob& J = K.jobs[n];
    J.out_off = off; J.w_off = w_off; J.fan_in = fan_in; J.col0 = col0; J.ncols_real = ncols_real; J.N = N;
    J.K = KS * h_slices; J.pre_kind = pre_kind; J.pre_off = pre_off;
    off += chain_step_bytes(s);
    ++n;
  };
  add(128, 1, 0, 0, 1, 0, BE_VIEWS, 0, B.grec_v, 0, 0, 0, 0, 1, L.rgb_w);
  add(TW, 0, 4, 1, 1, 0, BE_FEAT, 0, -1, L.views_w, L.W + L.in_ch_v, 0, TW, 0, 0);   // dZf: not stored (T trick)
  add(TW, 1, 8, 1, 1, 2, BE_TRUNK, L.D - 1, B.grec_l0 + (L.D - 1) * H_BYTES, L.feat_w, L.W, 0, TW, 2, L.alpha_w);
  int prev_wrote = 1;
  for (int l = L.D - 1; l >= 1; --l) {
    const bool is_skip = (L.skip >= 0 && l == L.skip + 1);
    if (is_skip && need_in) {
      add(64, 0, 8, prev_wrote, 0, 0, BE_ENC, 0, 0, L.pts_w[l], L.pts_in[l], 0, L.in_ch, 0, 0);
      C.steps[n - 1].keep_acc = 1;   // the next trunk step reads the same dZ (tensor memory of the OTHER buffer) again
      prev_wrote = 0;
      B.has_skip_enc = 1;
    }
    // the last dZ has a consumer inside the chain only when d/d gamma(x) is wanted
    add(TW, 0, 8, prev_wrote, (l > 1 || need_in) ? 1 : 0, 0, BE_TRUNK, l - 1, B.grec_l0 + (l - 1) * H_BYTES, L.pts_w[l],
        L.pts_in[l], is_skip ? L.in_ch : 0, TW, 0, 0);
    prev_wrote = 1;
  }
  if (need_in) add(64, 0, 8, prev_wrote, 0, 0, BE_ENC, 1, 0, L.pts_w[0], L.pts_in[0], 0, L.in_ch, 0, 0);
  C.nsteps = n;
  K.njobs = n;
  if (!chain_finish_plan(C, DgradRing::STAGE_BYTES)) C.nsteps = -1;   // rejected by the host entry point
}

// stream of a step: [pre16 block N x 16] [hidden blocks N x 64 ...]; element (n, k) = W[k][col0 + n]
__global__ void pack_bwd_weights_kernel(const float* __restrict__ params, __half* __restrict__ stream,
                                        BPackPlan plan) {
  const BPackJob& J = plan.jobs[blockIdx.y];
  const int n_pre = J.pre_kind ? J.N * 16 : 0;
  const int total = n_pre + J.N * J.K;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    float v = 0.f;
    if (e < n_pre) {
      const int g = e / (J.N * 8), rem = e - g * J.N * 8;
      const int n = rem >> 3, k = g * 8 + (rem & 7);
      if (J.pre_kind == 1) { if (k < 3) v = params[J.pre_off + (int64_t)k * J.N + n]; }
      else if (k == 3) v = params[J.pre_off + n];
    } else {
      const int r = e - n_pre;
      const int blk = r / (J.N * 64), rem = r - blk * J.N * 64;
      const int kk8 = rem / (J.N * 8), rem2 = rem - kk8 * J.N * 8;
      const int n = rem2 >> 3;
      const int k = blk * 64 + kk8 * 8 + (rem2 & 7);
      if (n < J.ncols_real) v = params[J.w_off + (int64_t)k * J.fan_in + J.col0 + n];
    }
    stream[J.out_off / 2 + e] = __float2half_rn(v);
  }
}

// max |g_out| as float bits (non-negative floats order like unsigned ints)
__global__ void absmax_kernel(const float* __restrict__ g, int64_t n, unsigned* __restrict__ out) {
  float m = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = fabsf(g[i]);
    if (v == v && v < INFINITY) m = fmaxf(m, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}
__device__ __forceinline__ float grad_scale_from(const unsigned* buf) {
  float m = __uint_as_float(*buf);
  if (!(m > 0.f)) return 1.f;
  float e = floorf(log2f(kGradTarget / m));
  e = fminf(fmaxf(e, -60.f), 60.f);
  return exp2f(e);
}

// one 128-bit read-only load.  (A `*reinterpret_cast<const uint4*>(p)` bound to a const reference parameter is read
// member by member: four 32-bit loads per lane, each touching 4x the sectors it needs -- head_grads, enc_bwd and
// ray_grad_reduce did that until round 2, ncu "L2 Theoretical Sectors Global Excessive" 75 %.)
__device__ __forceinline__ uint4 ldg128(const void* p) {
  uint4 q;
  asm("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(p));
  return q;
}
__device__ __forceinline__ void unpack8(const uint4 q, float (&v)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}

// shared-memory map of the dgrad kernel behind the ring and the activation tile
struct DgSmem {
  static constexpr int G16 = DgradRing::FREE0;                 // 2 x [128 x 16]: (g_r, g_g, g_b, g_sigma, 0...) * scale
  static constexpr int RED = G16 + 2 * CH_X16_BYTES;           // head-gradient partial sums of the prologue warps (fp32)
  static constexpr int RED_FLOATS = 4 * 256 + 8 * 384 + 128 * 4;
  static constexpr int TOTAL = RED + RED_FLOATS * 4;
};
// the head gradients (sigma head 256 -> 1, rgb head 128 -> 3) summed by the dgrad kernel's prologue warps
struct HeadFold {
  float* g_params;
  int h_last_off, hv_off;
  int64_t alpha_w, alpha_b, rgb_w, rgb_b;
};

// ------------------------------------------------------------------------------------------------
// stage 1: fused dgrad chain (one instance of the skeleton in mlp_chain.cuh)
// ------------------------------------------------------------------------------------------------
// ReLU mask words of `row` for the epilogue of one dgrad step: warpgroup g owns columns [64 rd + 32 g, +32) of every
// round rd; the forward kernel stores a layer's words as [g][row][rd], i.e. one aligned 16-byte (trunk) / 8-byte (views)
// load, contiguous over the rows of a warp.
__device__ __forceinline__ uint4 dgrad_mask_words(const uint8_t* frec, const FwdPlan& P, int epi, int p0, int row, int g) {
  uint4 m = make_uint4(0u, 0u, 0u, 0u);
  if (epi == BE_TRUNK) {
    m = *reinterpret_cast<const uint4*>(frec + P.rec_mask + ((p0 * 2 + g) * TROWS + row) * 16);
  } else if (epi == BE_VIEWS) {
    const uint2 v = *reinterpret_cast<const uint2*>(frec + P.rec_maskv + (g * TROWS + row) * 8);
    m.x = v.x; m.y = v.y;
  }
  return m;
}

__global__ void __launch_bounds__(CH_THREADS, 1)
mlp_tc_dgrad_kernel(const uint8_t* __restrict__ bstream, const float* __restrict__ g_out,
                    const unsigned* __restrict__ scale_buf, const uint8_t* __restrict__ fstash,
                    uint8_t* __restrict__ gstash, uint8_t* __restrict__ genc, int64_t M, ChainPlan C, FwdPlan P,
                    BwdPlan Bp, HeadFold HF) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ ChainBars B;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (M + TROWS - 1) / TROWS;
  // second 8-column group of both G16 tiles stays zero
  for (int i = tid; i < 2 * CH_X16_BYTES / 16; i += CH_THREADS)
    *reinterpret_cast<uint4*>(smem + DgSmem::G16 + i * 16) = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) chain_init_bars(B);
  if (warp == CH_WARP_MMA0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == CH_WARP_PRODUCER) {
    if (lane == 0) chain_producer<DgradRing>(B, smem, bstream, C, ntiles);
  } else if (warp >= CH_WARP_MMA0) {
    const ChainMap map = {0u, 0u, 0u, (uint32_t)DgSmem::G16, (uint32_t)CH_X16_BYTES};
    chain_mma_issuer<DgradRing>(B, smem, tmem_base, C, map, ntiles, (uint32_t)(warp - CH_WARP_MMA0));
  } else if (warp < CH_WG_EPI0) {
    // =========================== prologue warpgroup: loss-scaled head gradients of the NEXT tile ===========
    const int row = tid - CH_WG_PRO * 32;
    const float scale = grad_scale_from(scale_buf);
    // head gradients: these warps are idle once the tile's 2 KB operand is built, so they also sum
    //   dW_alpha[c] = sum_m g_sigma[m] H_{D-1}[m][c]      dW_rgb[k][c] = sum_m g_k[m] HV[m][c]      (+ the head biases)
    // over the tiles of this CTA -- fp32 g_out x the forward stash, as the separate head_grads kernel did.
    // Thread (lane = column group, warp = row quarter) for H_{D-1}; (tid & 15 = column group, tid >> 4 = 16-row block) for HV.
    float aa[8], ar[3][8], ab[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 8; ++i) { aa[i] = 0.f; ar[0][i] = 0.f; ar[1][i] = 0.f; ar[2][i] = 0.f; }
    const int cgV = row & 15, oV = row >> 4;
    int64_t it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int b = (int)(it & 1);
      const int64_t m = tile * TROWS + row;
      const float4 g = (m < M) ? reinterpret_cast<const float4*>(g_out)[m] : make_float4(0.f, 0.f, 0.f, 0.f);
      if (it >= 2) mbar_wait(&B.pro_free[b], (uint32_t)(((it >> 1) + 1) & 1));
      *reinterpret_cast<uint4*>(smem + DgSmem::G16 + b * CH_X16_BYTES + row * 16) =
          make_uint4(pack_h2(g.x * scale, g.y * scale), pack_h2(g.z * scale, g.w * scale), 0u, 0u);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&B.pro_ready[b]);
      {
        ab[0] += g.x; ab[1] += g.y; ab[2] += g.z; ab[3] += g.w;
        const uint8_t* rec = fstash + tile * (int64_t)P.rec_bytes;
        const int64_t m0 = tile * TROWS;
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += 8) {          // sigma head: rows [32 warp + r0, +8) of column group `lane`
          uint4 qv[8];
          float gs[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r = 32 * (warp - CH_WG_PRO) + r0 + j;
            qv[j] = ldg128(rec + HF.h_last_off + stash_off(32, r, lane));
            gs[j] = (m0 + r < M) ? __ldg(g_out + (m0 + r) * 4 + 3) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float v[8];
            unpack8(qv[j], v);
#pragma unroll
            for (int i = 0; i < 8; ++i) aa[i] = fmaf(gs[j], v[i], aa[i]);
          }
        }
#pragma unroll 1
        for (int r0 = 0; r0 < 16; r0 += 8) {          // rgb head: rows [16 oV + r0, +8) of column group cgV
          uint4 qv[8];
          float4 gr[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r = 16 * oV + r0 + j;
            qv[j] = ldg128(rec + HF.hv_off + stash_off(16, r, cgV));
            gr[j] = (m0 + r < M) ? __ldg(reinterpret_cast<const float4*>(g_out) + m0 + r) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float v[8];
            unpack8(qv[j], v);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              ar[0][i] = fmaf(gr[j].x, v[i], ar[0][i]);
              ar[1][i] = fmaf(gr[j].y, v[i], ar[1][i]);
              ar[2][i] = fmaf(gr[j].z, v[i], ar[2][i]);
            }
          }
        }
      }
    }
    {                          // partial sums of the 4 row quarters / 8 row blocks / 128 rows -> one atomic per gradient and CTA
      float* red = reinterpret_cast<float*>(smem + DgSmem::RED);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        red[(warp - CH_WG_PRO) * 256 + lane * 8 + i] = aa[i];
#pragma unroll
        for (int c = 0; c < 3; ++c) red[1024 + oV * 384 + c * 128 + cgV * 8 + i] = ar[c][i];
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) red[4096 + row * 4 + c] = ab[c];
      asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int col = 2 * row + k;
        atomicAdd(HF.g_params + HF.alpha_w + col, red[col] + red[256 + col] + red[512 + col] + red[768 + col]);
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int e = row + 128 * k;
        float t = 0.f;
#pragma unroll
        for (int o = 0; o < 8; ++o) t += red[1024 + o * 384 + e];
        atomicAdd(HF.g_params + HF.rgb_w + e, t);
      }
      if (row < 4) {
        float t = 0.f;
        for (int u = 0; u < 128; ++u) t += red[4096 + u * 4 + row];
        atomicAdd(HF.g_params + (row < 3 ? HF.rgb_b + row : HF.alpha_b), t);
      }
    }
  } else if (warp < CH_WARP_PRODUCER) {
    // =========================== epilogue warpgroups ===========================
    const int g = (warp - CH_WG_EPI0) >> 2;
    const int row = 32 * (warp & 3) + lane;
    const uint32_t tlane = (uint32_t)(32 * (warp & 3)) << 16;
    uint32_t d_phase = 0, acc_sel = 0;
    uint4 mnext = make_uint4(0u, 0u, 0u, 0u);
    if ((int64_t)blockIdx.x < ntiles)
      mnext = dgrad_mask_words(fstash + blockIdx.x * (int64_t)P.rec_bytes, P, C.steps[0].epi, C.steps[0].p0, row, g);
    const bool tracer = (warp == CH_WG_EPI0 + 4 * g);
    int64_t it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const uint8_t* frec = fstash + tile * (int64_t)P.rec_bytes;
      uint8_t* grec = gstash + tile * (int64_t)Bp.grec_bytes;
      for (int s = 0; s < C.nsteps; ++s) {
        const ChainStep S_ = C.steps[s];
        const int nround = S_.N >> 6;
        // the ReLU bits of my columns (one word per round) were requested one step ago (DRAM latency ~1 us)
        const uint4 mq = mnext;
        if (s + 1 < C.nsteps) mnext = dgrad_mask_words(frec, P, C.steps[s + 1].epi, C.steps[s + 1].p0, row, g);
        else if (tile + gridDim.x < ntiles)
          mnext = dgrad_mask_words(fstash + (tile + gridDim.x) * (int64_t)P.rec_bytes, P, C.steps[0].epi, C.steps[0].p0, row, g);
        mbar_wait(&B.d_ready[acc_sel], (d_phase >> acc_sel) & 1u); d_phase ^= (1u << acc_sel);
        tc_fence_after();
        if (tracer) chain_trace(C, it, s, 4 + 4 * g);
        const uint32_t tacc = tmem_base + tlane + acc_sel * 256u + 32u * g;
        uint32_t ra[32], rb2[32], pw[16];
        tmem_ld32(tacc, ra);
#pragma unroll 1
        for (int rd2 = 0; rd2 < nround; rd2 += 2) {
#pragma unroll
          for (int hb = 0; hb < 2; ++hb) {
            const int rd = rd2 + hb;
            if (rd < nround) {
              uint32_t (&r)[32] = hb ? rb2 : ra;
              const int col = 64 * rd + 32 * g;
              tmem_ld_wait();
              if (rd + 1 < nround) tmem_ld32(tacc + 64u * (rd + 1), hb ? ra : rb2);
              else {
                chain_release_acc(B, (int)acc_sel, lane);
                if (tracer) chain_trace(C, it, s, 5 + 4 * g);
              }
              pack_chunk<false>(r, pw);
              if (S_.epi == BE_TRUNK || S_.epi == BE_VIEWS) {   // ReLU gate of the producing layer
                const uint32_t m = rd2 == 0 ? (hb ? mq.y : mq.x) : (hb ? mq.w : mq.z);
                pw[0] = relu_mask_apply<0>(pw[0], m);   pw[1] = relu_mask_apply<1>(pw[1], m);
                pw[2] = relu_mask_apply<2>(pw[2], m);   pw[3] = relu_mask_apply<3>(pw[3], m);
                pw[4] = relu_mask_apply<4>(pw[4], m);   pw[5] = relu_mask_apply<5>(pw[5], m);
                pw[6] = relu_mask_apply<6>(pw[6], m);   pw[7] = relu_mask_apply<7>(pw[7], m);
                pw[8] = relu_mask_apply<8>(pw[8], m);   pw[9] = relu_mask_apply<9>(pw[9], m);
                pw[10] = relu_mask_apply<10>(pw[10], m); pw[11] = relu_mask_apply<11>(pw[11], m);
                pw[12] = relu_mask_apply<12>(pw[12], m); pw[13] = relu_mask_apply<13>(pw[13], m);
                pw[14] = relu_mask_apply<14>(pw[14], m); pw[15] = relu_mask_apply<15>(pw[15], m);
              }
              if (S_.epi != BE_ENC) {
                if (S_.writes_a) {
                  // A operand of the next step(s): packed over the first 16 columns of my own chunk of the buffer being drained
                  tmem_st16(tacc + 64u * rd, pw);
                  chain_publish_tmem(B, rd, lane);
                  if (tracer && (rd >> 1) == g) chain_trace(C, it, s, 6 + 4 * g + (rd & 1));
                }
                if (S_.epi != BE_FEAT) store_chunk_global(grec + S_.off0, S_.N >> 3, col, row, pw);
              } else {
                store_chunk_global(genc + (tile * 2 + S_.p0) * (int64_t)ENC_BYTES, 8, col, row, pw);
              }
            }
          }
        }
        if (S_.writes_a && nround < 4) {     // a 128-wide output: groups 2 and 3 do not exist, complete them empty
          __syncwarp();
          if (lane == 0) { mbar_arrive(&B.a_ready[2]); mbar_arrive(&B.a_ready[3]); }
        }
        if (!S_.keep_acc) acc_sel ^= 1u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == CH_WARP_MMA0) tmem_dealloc(tmem_base, 512);
}

// d/dx of the positional encoder from the two d/d gamma(x) blobs (loss-scaled fp16):  one thread per sample.
// sin / cos of 2^l x: exact sincosf at every third octave, the two octaves in between by angle doubling
// (sin 2a = 2 sin a cos a, cos 2a = 1 - 2 sin^2 a: <= 2 doublings, a few ulp) -- 12 instead of 30 sincosf per sample;
// the gradients these multiply are fp16 quantities.
__global__ void __launch_bounds__(128)
enc_bwd_kernel(const uint8_t* __restrict__ genc, const float* __restrict__ ray_batch, int ray_stride,
               const float* __restrict__ z, const unsigned* __restrict__ scale_buf, float* __restrict__ g_x,
               int64_t M, int S, int L_pts, int has_skip) {
  const int64_t tile = blockIdx.x;
  const int row = threadIdx.x;
  const int64_t m = tile * TROWS + row;
  if (m >= M) return;
  const float inv_scale = 1.f / grad_scale_from(scale_buf);
  float ge[64];
#pragma unroll
  for (int gq = 0; gq < 8; ++gq) {
    float v0[8], v1[8];
    if (has_skip) unpack8(ldg128(genc + (tile * 2 + 0) * (int64_t)ENC_BYTES + stash_off(8, row, gq)), v0);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v0[i] = 0.f;
    }
    unpack8(ldg128(genc + (tile * 2 + 1) * (int64_t)ENC_BYTES + stash_off(8, row, gq)), v1);
#pragma unroll
    for (int i = 0; i < 8; ++i) ge[gq * 8 + i] = v0[i] + v1[i];
  }
  const int64_t rr = m / S;
  const float* rb = ray_batch + rr * ray_stride;
  const float zz = z[m];
  float x[3];
  x[0] = __fadd_rn(rb[0], __fmul_rn(rb[3], zz));
  x[1] = __fadd_rn(rb[1], __fmul_rn(rb[4], zz));
  x[2] = __fadd_rn(rb[2], __fmul_rn(rb[5], zz));
  float gx[3] = {ge[0], ge[1], ge[2]};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float f = 1.f, sn = 0.f, cs = 1.f;
#pragma unroll
    for (int l = 0; l < 10; ++l) {
      if (l < L_pts) {
        if (l % 3 == 0) sincosf(x[c] * f, &sn, &cs);
        else { const float s2 = 2.f * sn * cs; cs = fmaf(-2.f * sn, sn, 1.f); sn = s2; }
        gx[c] += f * (cs * ge[3 + 6 * l + c] - sn * ge[3 + 6 * l + 3 + c]);
      }
      f *= 2.f;
    }
  }
  g_x[m * 3 + 0] = gx[0] * inv_scale;
  g_x[m * 3 + 1] = gx[1] * inv_scale;
  g_x[m * 3 + 2] = gx[2] * inv_scale;
}

// ------------------------------------------------------------------------------------------------
// stage 2: wgrad
// ------------------------------------------------------------------------------------------------
constexpr int WG_THREADS = 192;           // producer warp, MMA warp, 4 epilogue warps
constexpr int WG_STAGE_BYTES = 65536;     // 64 samples: dZ [64 x <=256] (32 KB) | X [64 x <=256] (32 KB)
constexpr int WG_NSTAGE = 3;
constexpr int MAX_WJOBS = 2 * kMaxDepth + 8;

struct WJob {
  int dz_off, nout;          // dZ blob inside the gstash record, its width (256 | 128)
  int x_off, kin, kin_real;  // X blob inside the forward stash record, padded / real width
  int64_t w_off;             // dW destination: g_params[w_off + n*fan_in + col0 + k]  (to_aux: aux[k*128 + n], transposed)
  int fan_in, col0, to_aux;
  int64_t b_off;             // bias-gradient destination (nout floats) or -1
};
constexpr int WG_MAX_UNITS = 6;           // units per CTA in the schedule table
struct WPlan {
  int njobs, nchunks, tiles_per_chunk;
  WJob jobs[MAX_WJOBS];
  // Units (job, chunk of tiles) differ in cost by 3x (a trunk layer streams 128 KB per tile, the viewdirs-encoding job
  // 40 KB), and the kernel is HBM-bound: the host deals them to the persistent CTAs longest-first onto the least loaded
  // CTA (LPT) and passes the table; a CTA's three roles walk the same list.
  uint8_t n_of[kNumSMs];
  uint16_t unit_of[kNumSMs][WG_MAX_UNITS];
};

static void make_wgrad_plan(const MlpLayout& L, const FwdPlan& P, const BwdPlan& B, int64_t ntiles, WPlan& Wp) {
  int n = 0;
  auto add = [&](int dz_off, int nout, int x_off, int kin, int kin_real, int64_t w_off, int fan_in, int col0,
                 int64_t b_off) {
    WJob& J = Wp.jobs[n++];
    J.dz_off = dz_off; J.nout = nout; J.x_off = x_off; J.kin = kin; J.kin_real = kin_real; J.w_off = w_off;
    J.fan_in = fan_in; J.col0 = col0; J.b_off = b_off; J.to_aux = 0;
  };
  // heaviest jobs first (units are dealt round-robin to the persistent CTAs)
  for (int l = L.D - 1; l >= 1; --l) {
    const bool is_skip = (L.skip >= 0 && l == L.skip + 1);
    add(B.grec_l0 + l * H_BYTES, TW, P.rec_h0 + (l - 1) * H_BYTES, TW, TW, L.pts_w[l], L.pts_in[l],
        is_skip ? L.in_ch : 0, L.pts_b[l]);
  }
  // T = dZv^T H_{D-1} into the aux buffer (+ sv = sum dZv straight into views bias gradient): see "T trick" above
  add(B.grec_v, 128, P.rec_h0 + (L.D - 1) * H_BYTES, TW, TW, 0, TW, 0, L.views_b);
  Wp.jobs[n - 1].to_aux = 1;
  add(B.grec_l0, TW, P.rec_encp, 64, L.in_ch, L.pts_w[0], L.pts_in[0], 0, L.pts_b[0]);
  if (L.skip >= 0 && L.skip + 1 < L.D)
    add(B.grec_l0 + (L.skip + 1) * H_BYTES, TW, P.rec_encp, 64, L.in_ch, L.pts_w[L.skip + 1], L.pts_in[L.skip + 1], 0, -1);
  add(B.grec_v, 128, P.rec_encv, 32, L.in_ch_v, L.views_w, L.W + L.in_ch_v, L.W, -1);
  Wp.njobs = n;
  // ~3 units per CTA (never more than the table holds): every unit pays a pipeline fill and a 65 536-atomic flush
  int64_t want_chunks = (3 * (int64_t)kNumSMs) / n;
  if (want_chunks > ntiles) want_chunks = ntiles;
  if (want_chunks < 1) want_chunks = 1;
  Wp.tiles_per_chunk = (int)((ntiles + want_chunks - 1) / want_chunks);
  Wp.nchunks = (int)((ntiles + Wp.tiles_per_chunk - 1) / Wp.tiles_per_chunk);
  // longest-processing-time-first assignment; cost of a unit = bytes it streams
  const int nunits = Wp.njobs * Wp.nchunks;
  static thread_local int order[MAX_WJOBS * 4 * kNumSMs];
  static thread_local int64_t cost[MAX_WJOBS * 4 * kNumSMs];
  for (int u = 0; u < nunits; ++u) {
    const WJob& J = Wp.jobs[u / Wp.nchunks];
    const int64_t t0 = (int64_t)(u % Wp.nchunks) * Wp.tiles_per_chunk;
    const int64_t t1 = (t0 + Wp.tiles_per_chunk < ntiles) ? t0 + Wp.tiles_per_chunk : ntiles;
    cost[u] = (t1 - t0) * (int64_t)(J.nout + J.kin) * 256 + 20000;   // + a per-unit flush / pipeline-fill term
    order[u] = u;
  }
  std::sort(order, order + nunits, [&](int a, int b) { return cost[a] != cost[b] ? cost[a] > cost[b] : a < b; });
  int64_t load[kNumSMs];
  const int ncta = nunits < kNumSMs ? nunits : kNumSMs;
  for (int c = 0; c < kNumSMs; ++c) { load[c] = 0; Wp.n_of[c] = 0; }
  for (int i = 0; i < nunits; ++i) {
    int best = -1;
    for (int c = 0; c < ncta; ++c)
      if (Wp.n_of[c] < WG_MAX_UNITS && (best < 0 || load[c] < load[best])) best = c;
    Wp.unit_of[best][Wp.n_of[best]++] = (uint16_t)order[i];
    load[best] += cost[order[i]];
  }
}

// 32 consecutive fp32 gradients of one output row, dst = address of value 0, LEAD = values before the first 16-byte
// boundary; only the first `nreal` values exist (the rest is K padding of the GEMM)
template <int LEAD>
__device__ __forceinline__ void wgrad_flush32(float* dst, const uint32_t (&r)[32], float inv_scale, int nreal) {
#pragma unroll
  for (int j = 0; j < LEAD; ++j)
    if (j < nreal) atomicAdd(dst + j, __uint_as_float(r[j]) * inv_scale);
#pragma unroll
  for (int j = LEAD; j + 4 <= 32; j += 4) {
    if (j + 4 <= nreal) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(r[j]) * inv_scale),
                   "f"(__uint_as_float(r[j + 1]) * inv_scale), "f"(__uint_as_float(r[j + 2]) * inv_scale),
                   "f"(__uint_as_float(r[j + 3]) * inv_scale)
                   : "memory");
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (j + i < nreal) atomicAdd(dst + j + i, __uint_as_float(r[j + i]) * inv_scale);
    }
  }
#pragma unroll
  for (int j = LEAD + 4 * ((32 - LEAD) / 4); j < 32; ++j)
    if (j < nreal) atomicAdd(dst + j, __uint_as_float(r[j]) * inv_scale);
}

__global__ void __launch_bounds__(WG_THREADS, 1)
mlp_tc_wgrad_kernel(const uint8_t* __restrict__ fstash, const uint8_t* __restrict__ gstash,
                    const unsigned* __restrict__ scale_buf, float* __restrict__ g_params, float* __restrict__ aux,
                    int64_t ntiles, int frec_bytes, int grec_bytes, WPlan Wp) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full_bar[WG_NSTAGE], empty_bar[WG_NSTAGE], acc_ready, acc_free;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < WG_NSTAGE; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1 + 4); }
    mbar_init(&acc_ready, 1);
    mbar_init(&acc_free, 128);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int my_units = Wp.n_of[blockIdx.x];

  if (warp == 0) {
    // ---- producer: a 64-sample half of a stash blob is one contiguous run (see stash_off) = the operand image
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int ui = 0; ui < my_units; ++ui) {
        const int u = Wp.unit_of[blockIdx.x][ui];
        const WJob& J = Wp.jobs[u / Wp.nchunks];
        const int64_t t0 = (int64_t)(u % Wp.nchunks) * Wp.tiles_per_chunk;
        const int64_t t1 = (t0 + Wp.tiles_per_chunk < ntiles) ? t0 + Wp.tiles_per_chunk : ntiles;
        const uint32_t abytes = (uint32_t)(J.nout / 8) * 1024u, bbytes = (uint32_t)(J.kin / 8) * 1024u;
        for (int64_t tile = t0; tile < t1; ++tile) {
          const uint8_t* asrc = gstash + tile * (int64_t)grec_bytes + J.dz_off;
          const uint8_t* bsrc = fstash + tile * (int64_t)frec_bytes + J.x_off;
          for (int half = 0; half < 2; ++half) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], abytes + bbytes);
            uint8_t* sa = smem + stage * WG_STAGE_BYTES;
            bulk_g2s(sa, asrc + half * abytes, abytes, &full_bar[stage]);
            bulk_g2s(sa + 32768, bsrc + half * bbytes, bbytes, &full_bar[stage]);
            if (++stage == WG_NSTAGE) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer (converged warp, one elected lane issues): per 64-sample stage 4 K=16 MMAs per output half,
    // both operands MN-major views of the staged blobs (16 samples = 256 bytes further per MMA)
    int stage = 0; uint32_t phase = 0, free_phase = 0;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t hi = desc_hi(1024u);
    bool first_unit = true;
    for (int ui = 0; ui < my_units; ++ui) {
      const int u = Wp.unit_of[blockIdx.x][ui];
      const WJob& J = Wp.jobs[u / Wp.nchunks];
      const int64_t t0 = (int64_t)(u % Wp.nchunks) * Wp.tiles_per_chunk;
      const int64_t t1 = (t0 + Wp.tiles_per_chunk < ntiles) ? t0 + Wp.tiles_per_chunk : ntiles;
      const uint32_t idesc = make_idesc_f16(128, J.kin, 1, 1);
      const int nhalf = J.nout / 128;
      if (!first_unit) { mbar_wait(&acc_free, free_phase); free_phase ^= 1; }
      first_unit = false;
      uint32_t acc = 0;
      for (int64_t it = 0; it < 2 * (t1 - t0); ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = desc_lo(sbase + stage * WG_STAGE_BYTES, 128u);
          const uint32_t b_lo = desc_lo(sbase + stage * WG_STAGE_BYTES + 32768, 128u);
          mma_ss<4>(tmem_base, a_lo, b_lo, hi, 16u, 16u, idesc, acc);
          if (nhalf == 2) mma_ss<4>(tmem_base + 256u, a_lo + 1024u, b_lo, hi, 16u, 16u, idesc, acc);
          umma_commit(&empty_bar[stage]);
        }
        acc = 1;
        if (++stage == WG_NSTAGE) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&acc_ready);
    }
  } else {
    // ---- epilogue warps: bias column sums from the dZ stage while the MMAs run, then the accumulator flush
    const int et = tid - 64;                 // 0..127
    const int row = 32 * (warp & 3) + lane;  // TMEM lane = output unit within the 128-unit half
    const uint32_t tlane = (uint32_t)(32 * (warp & 3)) << 16;
    const int q = et >> 3, r8 = et & 7;
    const float inv_scale = 1.f / grad_scale_from(scale_buf);
    int stage = 0; uint32_t phase = 0, ready_phase = 0;
    for (int ui = 0; ui < my_units; ++ui) {
      const int u = Wp.unit_of[blockIdx.x][ui];
      const WJob& J = Wp.jobs[u / Wp.nchunks];
      const int64_t t0 = (int64_t)(u % Wp.nchunks) * Wp.tiles_per_chunk;
      const int64_t t1 = (t0 + Wp.tiles_per_chunk < ntiles) ? t0 + Wp.tiles_per_chunk : ntiles;
      const int ngrp = J.nout / 8;           // 32 or 16 column groups
      float bs[2][8];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int i = 0; i < 8; ++i) bs[a][i] = 0.f;
      for (int64_t it = 0; it < 2 * (t1 - t0); ++it) {
        mbar_wait(&full_bar[stage], phase);
        if (J.b_off >= 0) {
          const uint8_t* sa = smem + stage * WG_STAGE_BYTES;
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const int gi = q + 16 * a;
            if (gi < ngrp) {
              // 8 rows x 8 columns: partial sums of 8 values in packed fp16, then one conversion to fp32
              uint4 acc4 = *reinterpret_cast<const uint4*>(sa + gi * 1024 + r8 * 16);
              __half2* ah = reinterpret_cast<__half2*>(&acc4);
#pragma unroll
              for (int rr = 1; rr < 8; ++rr) {
                const uint4 v4 = *reinterpret_cast<const uint4*>(sa + gi * 1024 + (r8 + 8 * rr) * 16);
                const __half2* vh = reinterpret_cast<const __half2*>(&v4);
#pragma unroll
                for (int i = 0; i < 4; ++i) ah[i] = __hadd2(ah[i], vh[i]);
              }
              float v[8];
              unpack8(acc4, v);
#pragma unroll
              for (int i = 0; i < 8; ++i) bs[a][i] += v[i];
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == WG_NSTAGE) { stage = 0; phase ^= 1; }
      }
      if (J.b_off >= 0) {
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const int gi = q + 16 * a;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float v = bs[a][i];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            if (r8 == 0 && gi < ngrp) atomicAdd(g_params + J.b_off + gi * 8 + i, v * inv_scale);
          }
        }
      }
      mbar_wait(&acc_ready, ready_phase); ready_phase ^= 1;
      tc_fence_after();
      for (int h = 0; h < J.nout / 128; ++h) {
        float* dst = (J.to_aux ? aux : g_params + J.w_off) + (int64_t)(h * 128 + row) * J.fan_in + J.col0;
        // a lane owns one output row: its 32 values of a column block are contiguous in dW and leave as four-float
        // reductions on 16-byte boundaries instead of 32 scalar ones (a quarter of the L2 atomic operations: ncu counted
        // 32 sectors per scalar REDG, 20 M sector-atomics per launch)
        // (rows of a 319- or 283-column weight start at every alignment: the lane's first 0..3 values go as scalars)
        const int lead = (4 - (int)((reinterpret_cast<uintptr_t>(dst) >> 2) & 3)) & 3;
#pragma unroll 1
        for (int cc = 0; cc < J.kin; cc += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + tlane + h * 256 + cc, r);
          tmem_ld_wait();
          if (J.to_aux) {
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(aux + (cc + j) * 128 + row, __uint_as_float(r[j]) * inv_scale);   // T^T[k][n]: coalesced
          } else {
            const int nreal = J.kin_real - cc;      // values of this block that exist (>= 32: all)
            switch (lead) {
              case 0: wgrad_flush32<0>(dst + cc, r, inv_scale, nreal); break;
              case 1: wgrad_flush32<1>(dst + cc, r, inv_scale, nreal); break;
              case 2: wgrad_flush32<2>(dst + cc, r, inv_scale, nreal); break;
              default: wgrad_flush32<3>(dst + cc, r, inv_scale, nreal); break;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_free);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// stage 3b: per-ray reduction  d/d(o), d/d(d) from per-sample d/dx;  d/d(viewdirs) from sum_s dZv
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
ray_grad_reduce_kernel(const float* __restrict__ g_x, const float* __restrict__ z,
                       const uint8_t* __restrict__ gstash, int grec_bytes, int grec_v,
                       const float* __restrict__ params, int64_t views_w, int fan_in_v, int in_ch_v, int L_views,
                       const float* __restrict__ ray_batch, int ray_stride, const unsigned* __restrict__ scale_buf,
                       float* __restrict__ g_rb, int64_t R, int S) {
  __shared__ float s_sum[4][128];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * 4 + w;
  if (r >= R) return;
  const float inv_scale = 1.f / grad_scale_from(scale_buf);
  float go[3] = {0, 0, 0}, gd[3] = {0, 0, 0};
  for (int s = lane; s < S; s += 32) {
    const int64_t m = r * S + s;
    const float zz = z[m];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = g_x[m * 3 + c];
      go[c] += v;
      gd[c] += v * zz;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) { go[c] = warp_sum(go[c]); gd[c] = warp_sum(gd[c]); }
  // sum of the views-layer dZ over the samples of this ray: four column groups per pass, all of their loads in flight
  // together (one DRAM round trip per pass instead of one per column group)
#pragma unroll 1
  for (int g0 = 0; g0 < 16; g0 += 4) {
    float a[4][8];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[q][i] = 0.f;
#pragma unroll 1
    for (int s0 = lane; s0 < S; s0 += 64) {
      uint4 qv[2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int s = s0 + 32 * u;
        const int64_t m = r * S + (s < S ? s : s0);
        const int64_t tile = m / TROWS;
        const int row = (int)(m - tile * TROWS);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          qv[u][q] = ldg128(gstash + tile * (int64_t)grec_bytes + grec_v + stash_off(16, row, g0 + q));
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (s0 + 32 * u < S) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v[8];
            unpack8(qv[u][q], v);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[q][i] += v[i];
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float t = warp_sum(a[q][i]);
        if (lane == i) s_sum[w][(g0 + q) * 8 + i] = t;
      }
  }
  __syncwarp();
  float ge = 0.f;   // lane q < in_ch_v: d/d gamma(v)[q]
  if (lane < in_ch_v) {
    for (int j = 0; j < 128; ++j) ge = fmaf(s_sum[w][j], params[views_w + (int64_t)j * fan_in_v + TW + lane], ge);
    ge *= inv_scale;
  }
  const float* rb = ray_batch + r * ray_stride;
  float gv = 0.f;
  {
    const int c = lane < 3 ? lane : 0;
    const float x = rb[8 + c];
    gv = __shfl_sync(0xffffffffu, ge, c);
    float f = 1.f;
    for (int l = 0; l < L_views; ++l) {
      float sn, cs;
      sincosf(x * f, &sn, &cs);
      const float gs = __shfl_sync(0xffffffffu, ge, 3 + 6 * l + c);
      const float gc = __shfl_sync(0xffffffffu, ge, 3 + 6 * l + 3 + c);
      gv += f * (cs * gs - sn * gc);
      f *= 2.f;
    }
  }
  float* o = g_rb + r * 11;
  if (lane == 0) {
    o[0] = go[0]; o[1] = go[1]; o[2] = go[2]; o[3] = gd[0]; o[4] = gd[1]; o[5] = gd[2]; o[6] = 0.f; o[7] = 0.f;
  }
  if (lane < 3) o[8 + lane] = gv;
}

// ------------------------------------------------------------------------------------------------
// stage 2b: the two small products of the "T trick" (once per call).  aux = T^T [256][128] with T = dZv^T H_{D-1}
// (the wgrad flush of that job writes the transposed orientation: lanes = consecutive n, coalesced atomics);
// sv = views bias gradient (already final in g_params).  Weights enter fp16-rounded, as the chain kernels used them.
//   dW_f[j][k] = sum_n Wv[n][j] T^T[k][n]            db_f[j] = sum_n Wv[n][j] sv[n]
//   dW_v[n][j] = sum_k T^T[k][n] Wf[j][k] + sv[n] bf[j]      (j = feature columns 0..255 of the views weight)
// One thread per output, no staging: in both products the lanes of a warp read CONSECUTIVE elements of one operand
// (Wv along j / T^T along n) and the SAME element of the other (a broadcast), 2.4 resident blocks per SM.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float h16r(float v) { return __half2float(__float2half_rn(v)); }
__global__ void __launch_bounds__(256)
fv_post_kernel(const float* __restrict__ aux, const float* __restrict__ params, float* __restrict__ g_params,
               int64_t views_w, int fan_in_v, int64_t views_b, int64_t feat_w, int64_t feat_b) {
  const float* __restrict__ Tt = aux;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < TW * TW) {                                   // dW_f[j][k]: lanes along j
    const int k = e / TW, j = e - k * TW;
    const float* wv = params + views_w + j;
    const float* tk = Tt + k * 128;
    float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
    for (int n = 0; n < 128; n += 2) {
      a0 = fmaf(h16r(wv[(int64_t)n * fan_in_v]), tk[n], a0);
      a1 = fmaf(h16r(wv[(int64_t)(n + 1) * fan_in_v]), tk[n + 1], a1);
    }
    g_params[feat_w + (int64_t)j * TW + k] = a0 + a1;
  } else if (e < TW * TW + 128 * TW) {                 // dW_v[n][j]: lanes along n
    const int q = e - TW * TW, j = q >> 7, n = q & 127;
    const float* wf = params + feat_w + (int64_t)j * TW;
    float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
    for (int k = 0; k < TW; k += 2) {
      a0 = fmaf(Tt[k * 128 + n], h16r(wf[k]), a0);
      a1 = fmaf(Tt[(k + 1) * 128 + n], h16r(wf[k + 1]), a1);
    }
    g_params[views_w + (int64_t)n * fan_in_v + j] = fmaf(g_params[views_b + n], h16r(params[feat_b + j]), a0 + a1);
  } else if (e < TW * TW + 128 * TW + TW) {            // db_f[j]
    const int j = e - TW * TW - 128 * TW;
    float acc = 0.f;
    for (int n = 0; n < 128; ++n) acc = fmaf(h16r(params[views_w + (int64_t)n * fan_in_v + j]), g_params[views_b + n], acc);
    g_params[feat_b + j] = acc;
  }
}

}  // namespace lush

using namespace lush;

int lush_tc_gather_head_params(const MlpLayout& L, const float* params, float* dst, cudaStream_t st);  // mlp_tc.cu

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

struct BwdWorkspace {
  size_t bstream, head, scale, gx, genc, gstash, total;
};
static BwdWorkspace bwd_ws_layout(const MlpLayout& L, int64_t M, bool need_in) {
  FwdPlan P; make_fwd_plan(L, P);
  ChainPlan C; BwdPlan B; BPackPlan K; make_bwd_chain(L, need_in, C, B, K);
  const int64_t ntiles = (M + TROWS - 1) / TROWS;
  BwdWorkspace w;
  size_t off = 0;
  w.bstream = off; off = align256(off + (size_t)C.stream_bytes);
  w.head = off; off = align256(off + (size_t)P.nparam_floats * 4);
  w.scale = off; off = align256(off + 256);          // [loss-scale word (256 B) | aux: T 128 x 256 fp32] zeroed by ONE memset
  off = align256(off + (size_t)128 * TW * 4);
  w.gx = off; off = align256(off + (size_t)M * 3 * 4);
  w.genc = off; off = align256(off + (size_t)ntiles * 2 * ENC_BYTES);
  w.gstash = off; off = align256(off + (size_t)ntiles * B.grec_bytes);
  w.total = off;
  return w;
}

size_t lush_mlp_tc_bwd_workspace_bytes(const MlpLayout& L, int64_t R, int S) {
  return bwd_ws_layout(L, R * S, true).total;
}

int lush_mlp_tc_bwd(const MlpLayout& L, const lush_mlp_desc* desc, const float* params, const float* ray_batch,
                    int ray_stride, const float* z, const float* g_out, const void* stash, float* g_params,
                    float* g_ray_batch, void* workspace, size_t workspace_bytes, int64_t R, int S, cudaStream_t st) {
  const int64_t M = R * S;
  const bool need_in = g_ray_batch != nullptr;
  FwdPlan P; make_fwd_plan(L, P);
  ChainPlan C; BwdPlan B; BPackPlan K; make_bwd_chain(L, need_in, C, B, K);
  LUSH_REQUIRE(C.nsteps > 0, "mlp_bwd(tc): a weight block of this topology does not fit the ring stage");
  const BwdWorkspace w = bwd_ws_layout(L, M, true);
  if (!workspace || workspace_bytes < w.total) {
    set_error("mlp_bwd(tc): workspace too small (%zu < %zu)", workspace_bytes, w.total);
    return LUSH_EWORKSPACE;
  }
  LUSH_REQUIRE(((uintptr_t)workspace & 255) == 0 && ((uintptr_t)stash & 15) == 0 && ((uintptr_t)g_out & 15) == 0,
               "mlp_bwd(tc): unaligned buffers");
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  __half* bstream = reinterpret_cast<__half*>(ws + w.bstream);
  float* head = reinterpret_cast<float*>(ws + w.head);
  unsigned* scale_buf = reinterpret_cast<unsigned*>(ws + w.scale);
  float* aux = reinterpret_cast<float*>(ws + w.scale + 256);
  float* gx = reinterpret_cast<float*>(ws + w.gx);
  uint8_t* genc = ws + w.genc;
  uint8_t* gstash = ws + w.gstash;
  const uint8_t* fstash = reinterpret_cast<const uint8_t*>(stash);
  const int64_t ntiles = (M + TROWS - 1) / TROWS;

  int rc = check_cuda(cudaMemsetAsync(scale_buf, 0, 256 + (size_t)128 * TW * 4, st), "mlp_bwd(tc): memset");   // scale word + T^T
  if (rc) return rc;
  absmax_kernel<<<(unsigned)(M * 4 / 1024 < 1 ? 1 : (M * 4 / 1024 > 1184 ? 1184 : M * 4 / 1024)), 256, 0, st>>>(g_out, M * 4, scale_buf);
  LUSH_LAUNCH_CHECK("absmax");
  pack_bwd_weights_kernel<<<dim3(64, K.njobs), 256, 0, st>>>(params, bstream, K);
  LUSH_LAUNCH_CHECK("pack_bwd_weights");
  rc = lush_tc_gather_head_params(L, params, head, st);
  if (rc) return rc;
  {
    const size_t smem = DgSmem::TOTAL;
    rc = check_cuda(cudaFuncSetAttribute(mlp_tc_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                    "mlp_bwd(tc): dgrad smem attr");
    if (rc) return rc;
    int grid = (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
    const HeadFold HF = {g_params, P.rec_h0 + (L.D - 1) * H_BYTES, P.rec_hv, L.alpha_w, L.alpha_b, L.rgb_w, L.rgb_b};
    mlp_tc_dgrad_kernel<<<grid, CH_THREADS, smem, st>>>((const uint8_t*)bstream, g_out, scale_buf, fstash, gstash, genc,
                                                       M, C, P, B, HF);
    LUSH_LAUNCH_CHECK("mlp_tc_dgrad");
  }
  {
    WPlan Wp; make_wgrad_plan(L, P, B, ntiles, Wp);
    const size_t smem = (size_t)WG_NSTAGE * WG_STAGE_BYTES;
    rc = check_cuda(cudaFuncSetAttribute(mlp_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                    "mlp_bwd(tc): wgrad smem attr");
    if (rc) return rc;
    const int nunits = Wp.njobs * Wp.nchunks;
    int grid = nunits < kNumSMs ? nunits : kNumSMs;
    mlp_tc_wgrad_kernel<<<grid, WG_THREADS, smem, st>>>(fstash, gstash, scale_buf, g_params, aux, ntiles, P.rec_bytes,
                                                       B.grec_bytes, Wp);
    LUSH_LAUNCH_CHECK("mlp_tc_wgrad");
    fv_post_kernel<<<(TW * TW + 128 * TW + TW + 255) / 256, 256, 0, st>>>(aux, params, g_params, L.views_w, L.W + L.in_ch_v,
                                                                         L.views_b, L.feat_w, L.feat_b);
    LUSH_LAUNCH_CHECK("fv_post");
  }
  if (need_in) {
    enc_bwd_kernel<<<(unsigned)ntiles, 128, 0, st>>>(genc, ray_batch, ray_stride, z, scale_buf, gx, M, S, desc->L_pts,
                                                     B.has_skip_enc);
    LUSH_LAUNCH_CHECK("enc_bwd");
    ray_grad_reduce_kernel<<<(unsigned)((R + 3) / 4), 128, 0, st>>>(gx, z, gstash, B.grec_bytes, B.grec_v, params,
                                                                   L.views_w, L.W + L.in_ch_v, L.in_ch_v,
                                                                   desc->L_views, ray_batch, ray_stride, scale_buf,
                                                                   g_ray_batch, R, S);
    LUSH_LAUNCH_CHECK("ray_grad_reduce");
  }
  return 0;
}

extern "C" int lush_debug_read_chain_trace_bwd(long long* host_out, int n) {   // dgrad kernel's trace (this translation unit)
  LUSH_REQUIRE(host_out && n > 0 && n <= CH_MAX_STEPS * 16, "read_chain_trace_bwd: bad arguments");
  return check_cuda(cudaMemcpyFromSymbol(host_out, g_chain_trace, sizeof(long long) * n), "read_chain_trace_bwd");
}
