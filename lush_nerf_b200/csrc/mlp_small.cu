// Fused exact-fp32 forward / backward of SMALL networks over FEW rows (the 4x128 noise-estimation MLP on
// N_rand pixels; reference helpers:456-512, models/lushnerf.py:268-293): one launch runs the positional encoder
// and every layer of a tile of rows with the activations in shared memory, instead of ~20 (forward) / ~25
// (backward) latency-bound GEMM launches of csrc/mlp_simt.cu.  Same arithmetic (fp32 FFMA, fp32 accumulate),
// same stash layout (SimtStash), so either backward can consume either forward's stash.
//
// Forward: CTA = 128 threads, TM rows.  Thread n owns output unit n of the current layer for all TM rows
// (TM accumulators); the layer input sits in shared memory and is read as broadcast float4s.
// Backward (d/d inputs not supported: the noise MLP is fed detached points): per layer, with dZ kept TRANSPOSED
// in shared memory (dZt[n][r]):   dW[n][k] += sum_r dZt[n][r] X[r][k]   (thread = k, X column in registers,
// one fp32 atomic per (n,k) and CTA),   db[n] += sum_r dZt[n][r],   g_prev[r][k] = sum_n dZt[n][r] W[n][k]
// masked by X[r][k] > 0 (thread = k, coalesced weight reads).
#include "mlp_common.cuh"

namespace lush {

constexpr int SM_THREADS = 128;
// rows per CTA below SM_BIG_ROWS rows (training: N_rand = 1024 rows).  Measured on B200 (tools/jobs/r2_run40.sh, r2_run41.sh;
// forward + backward per step): 8 / 16 -> 0.160 ms, 4 / 8 -> 0.126, 2 / 4 -> 0.142, 4 / 4 -> 0.116: with one warp per SM
// sub-partition nothing hides a CTA's latencies, so more, smaller CTAs win until the per-CTA weight walk dominates.
#ifndef LUSH_SM_FWD_TM
#define LUSH_SM_FWD_TM 4
#endif
#ifndef LUSH_SM_BWD_TM
#define LUSH_SM_BWD_TM 4
#endif
constexpr int SM_FWD_TM = LUSH_SM_FWD_TM;
constexpr int SM_FWD_TM_BIG = 32;  // rows per CTA from SM_BIG_ROWS rows on (rendering: a 32 768-ray chunk): every CTA walks
constexpr int64_t SM_BIG_ROWS = 4096;   // all 229 KB of weights once per row tile, so 4x the rows = 1/4 of the L2 traffic;
                                   // same FMA order per output, bit-identical results
constexpr int SM_BWD_TM = LUSH_SM_BWD_TM;
constexpr int SM_XLD = 192;        // row stride of the layer-input tile (>= W + in_ch, W <= 128, in_ch <= 63)

struct SmallStash {                // same arrays as SimtStash in mlp_simt.cu
  float *enc_pts, *enc_views, *h[kMaxDepth], *feat, *hv;
};

// out[r][n] = act(b[n] + sum_k X[r][k] W[n][wc + k]) over one or two input segments; result to shared / global
template <int TM>
__device__ __forceinline__ void small_dense(int t, int N, const float* __restrict__ x0, int ld0, int K0, int wc0,
                                            const float* __restrict__ x1, int ld1, int K1, int wc1,
                                            const float* __restrict__ Wt, int fan, const float* __restrict__ bias,
                                            bool relu, float* __restrict__ dst_s, int ldd, float* __restrict__ dst_g,
                                            int ldg, int64_t m0, int64_t M) {
  float acc[TM];
  if (t < N) {
    const float b = bias[t];
#pragma unroll
    for (int r = 0; r < TM; ++r) acc[r] = b;
    const float* wrow = Wt + (int64_t)t * fan;
    // every thread walks its own weight row: request all of its cache lines up front, so the walk pays one L2 round
    // trip instead of one per 128 bytes
    for (int k = 0; k < fan; k += 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(wrow + k));
#pragma unroll 1
    for (int seg = 0; seg < 2; ++seg) {
      const float* x = seg ? x1 : x0;
      const int ld = seg ? ld1 : ld0, K = seg ? K1 : K0, wc = seg ? wc1 : wc0;
      if (x == nullptr) continue;
      // every thread streams its own weight row: 16-byte loads when the row segment is aligned (4x fewer L1 requests)
      const bool vec = ((fan | wc) & 3) == 0 && (K & 3) == 0 && (reinterpret_cast<uintptr_t>(Wt) & 15) == 0;
      // The walk is bound by L2 latency, so NB quads of weights are requested one batch ahead of the FMAs that use them
      // (two explicit register buffers: left to itself the compiler funnelled every load of the unrolled loop through
      // ONE register quad, i.e. one L2 round trip per four weights -- 65 us for a 4x128 MLP on 1024 rows).
      constexpr int NB = TM >= 32 ? 4 : 8;
      const int nq = (K + 3) >> 2;          // shared rows are zero-padded to a multiple of 4
      auto load_batch = [&](int q0, float (&w)[NB][4]) {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int kq = (q0 + j) * 4;
          if (vec && kq < K) {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(wrow + wc + kq));
            w[j][0] = w4.x; w[j][1] = w4.y; w[j][2] = w4.z; w[j][3] = w4.w;
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) w[j][i] = (!vec && kq + i < K) ? __ldg(wrow + wc + kq + i) : 0.f;
          }
        }
      };
      auto fma_batch = [&](int q0, const float (&w)[NB][4]) {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int kq = (q0 + j) * 4;
          if (kq < K) {
#pragma unroll
            for (int r = 0; r < TM; ++r) {
              const float4 a = *reinterpret_cast<const float4*>(x + r * ld + kq);
              acc[r] = fmaf(a.x, w[j][0], acc[r]);
              acc[r] = fmaf(a.y, w[j][1], acc[r]);
              acc[r] = fmaf(a.z, w[j][2], acc[r]);
              acc[r] = fmaf(a.w, w[j][3], acc[r]);
            }
          }
        }
      };
      float wa[NB][4], wb[NB][4];
      load_batch(0, wa);
#pragma unroll 1
      for (int q0 = 0; q0 < nq; q0 += 2 * NB) {
        load_batch(q0 + NB, wb);
        fma_batch(q0, wa);
        load_batch(q0 + 2 * NB, wa);
        fma_batch(q0 + NB, wb);
      }
    }
  }
  if (t < N) {
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      const float v = relu ? fmaxf(acc[r], 0.f) : acc[r];
      if (dst_s) dst_s[r * ldd + t] = v;
      if (dst_g && m0 + r < M) dst_g[(m0 + r) * ldg + t] = v;
    }
  }
  __syncthreads();
}

template <int TM>
__global__ void __launch_bounds__(SM_THREADS)
small_mlp_fwd_kernel(MlpLayout L, const float* __restrict__ params, const float* __restrict__ ray_batch,
                     int ray_stride, const float* __restrict__ z, float* __restrict__ out, SmallStash s, int64_t M,
                     int S, int Lp, int Lv) {
  __shared__ __align__(16) float enc[TM][64];
  __shared__ __align__(16) float encv[TM][32];
  __shared__ __align__(16) float bufa[TM][128];
  __shared__ __align__(16) float bufb[TM][128];
  const int t = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * TM;
  const int W = L.W, W2 = L.W / 2;
  for (int i = t; i < TM * 64; i += SM_THREADS) (&enc[0][0])[i] = 0.f;
  for (int i = t; i < TM * 32; i += SM_THREADS) (&encv[0][0])[i] = 0.f;
  __syncthreads();
  for (int idx = t; idx < TM * 3; idx += SM_THREADS) {
    const int r = idx / 3, c = idx - r * 3;
    const int64_t m = m0 + r;
    if (m < M) {
      const float* rb = ray_batch + (m / S) * ray_stride;
      const float x = __fadd_rn(rb[c], __fmul_rn(rb[3 + c], z[m]));   // o + d*z, unfused like the reference
      const float v = rb[8 + c];
      enc[r][c] = x; encv[r][c] = v;
      float f = 1.f;
      for (int l = 0; l < Lp; ++l) {
        float sn, cs; sincosf(x * f, &sn, &cs);
        enc[r][3 + 6 * l + c] = sn; enc[r][3 + 6 * l + 3 + c] = cs; f *= 2.f;
      }
      f = 1.f;
      for (int l = 0; l < Lv; ++l) {
        float sn, cs; sincosf(v * f, &sn, &cs);
        encv[r][3 + 6 * l + c] = sn; encv[r][3 + 6 * l + 3 + c] = cs; f *= 2.f;
      }
    }
  }
  __syncthreads();
  if (s.enc_pts)
    for (int idx = t; idx < TM * L.in_ch; idx += SM_THREADS) {
      const int r = idx / L.in_ch, k = idx - r * L.in_ch;
      if (m0 + r < M) s.enc_pts[(m0 + r) * L.in_ch + k] = enc[r][k];
    }
  if (s.enc_views)
    for (int idx = t; idx < TM * L.in_ch_v; idx += SM_THREADS) {
      const int r = idx / L.in_ch_v, k = idx - r * L.in_ch_v;
      if (m0 + r < M) s.enc_views[((m0 + r) / S) * L.in_ch_v + k] = encv[r][k];   // same value from every sample of a ray
    }
  float* cur = nullptr;
  float* nxt = &bufa[0][0];
  for (int i = 0; i < L.D; ++i) {
    const bool skip = (L.skip >= 0 && i == L.skip + 1);
    const float* x0 = (i == 0 || skip) ? &enc[0][0] : cur;
    const int ld0 = (i == 0 || skip) ? 64 : 128, K0 = (i == 0 || skip) ? L.in_ch : W;
    const float* x1 = skip ? cur : nullptr;
    small_dense<TM>(t, W, x0, ld0, K0, 0, x1, 128, W, L.in_ch, params + L.pts_w[i], L.pts_in[i], params + L.pts_b[i],
                    true, nxt, 128, s.h[i], W, m0, M);
    cur = nxt;
    nxt = (cur == &bufa[0][0]) ? &bufb[0][0] : &bufa[0][0];
  }
  if (L.out_mode == 0)   // sigma head reads the last trunk activation
    small_dense<TM>(t, 1, cur, 128, W, 0, nullptr, 0, 0, 0, params + L.alpha_w, W, params + L.alpha_b, false, nullptr, 0,
                    out + 3, L.out_ch, m0, M);
  small_dense<TM>(t, W, cur, 128, W, 0, nullptr, 0, 0, 0, params + L.feat_w, W, params + L.feat_b, false, nxt, 128,
                  s.feat, W, m0, M);
  float* feat = nxt;
  float* hvb = cur;      // the last trunk activation is dead now
  small_dense<TM>(t, W2, feat, 128, W, 0, &encv[0][0], 32, L.in_ch_v, W, params + L.views_w, W + L.in_ch_v,
                  params + L.views_b, true, hvb, 128, s.hv, W2, m0, M);
  small_dense<TM>(t, 3, hvb, 128, W2, 0, nullptr, 0, 0, 0, params + L.rgb_w, W2, params + L.rgb_b, false, nullptr, 0, out,
                  L.out_ch, m0, M);
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// dW[n][wc + k] += sum_r dzt[n][r] X[r][k]  for k in [0,K);  thread = k (+128, ...), X column in registers
template <int TM>
__device__ __forceinline__ void small_wgrad(int t, int N, const float (*dzt)[TM], const float* __restrict__ x, int ld,
                                            int K, float* __restrict__ dW, int fan, int wc) {
  for (int k = t; k < K; k += SM_THREADS) {
    float xr[TM];
#pragma unroll
    for (int r = 0; r < TM; ++r) xr[r] = x[r * ld + k];
#pragma unroll 2
    for (int n = 0; n < N; ++n) {
      float sacc = 0.f;
#pragma unroll
      for (int r4 = 0; r4 < TM; r4 += 4) {
        const float4 d = *reinterpret_cast<const float4*>(&dzt[n][r4]);
        sacc = fmaf(d.x, xr[r4], sacc); sacc = fmaf(d.y, xr[r4 + 1], sacc);
        sacc = fmaf(d.z, xr[r4 + 2], sacc); sacc = fmaf(d.w, xr[r4 + 3], sacc);
      }
      atomicAdd(dW + (int64_t)n * fan + wc + k, sacc);
    }
  }
}
template <int TM>
__device__ __forceinline__ void small_bgrad(int t, int N, const float (*dzt)[TM], float* __restrict__ db) {
  if (t < N) {
    float sacc = 0.f;
#pragma unroll
    for (int r = 0; r < TM; ++r) sacc += dzt[t][r];
    atomicAdd(db + t, sacc);
  }
}
// gt[k][r] = (mask ? x_mask[r][k] > 0 : 1) * (sum_n dzt[n][r] W[n][wc + k] + extra[r] * wextra[k])  for k in [0,K), K <= 128
template <int TM>
__device__ __forceinline__ void small_dgrad(int t, int N, int K, const float (*dzt)[TM], const float* __restrict__ Wt,
                                            int fan, int wc, const float* __restrict__ xmask, int ldm,
                                            const float* __restrict__ extra, const float* __restrict__ wextra,
                                            float (*gt)[TM]) {
  if (t < K) {
    float acc[TM];
#pragma unroll
    for (int r = 0; r < TM; ++r) acc[r] = 0.f;
    // weights of 8 output units per batch, requested one batch ahead of their FMAs (explicit double buffer, see small_dense)
    const float* wcol = Wt + wc + t;
    auto load8 = [&](int n0, float (&w)[8]) {
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = (n0 + j < N) ? __ldg(wcol + (int64_t)(n0 + j) * fan) : 0.f;
    };
    auto fma8 = [&](int n0, const float (&w)[8]) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (n0 + j < N) {
#pragma unroll
          for (int r4 = 0; r4 < TM; r4 += 4) {
            const float4 d = *reinterpret_cast<const float4*>(&dzt[n0 + j][r4]);
            acc[r4] = fmaf(d.x, w[j], acc[r4]); acc[r4 + 1] = fmaf(d.y, w[j], acc[r4 + 1]);
            acc[r4 + 2] = fmaf(d.z, w[j], acc[r4 + 2]); acc[r4 + 3] = fmaf(d.w, w[j], acc[r4 + 3]);
          }
        }
      }
    };
    float wa[8], wb[8];
    load8(0, wa);
#pragma unroll 1
    for (int n0 = 0; n0 < N; n0 += 16) {
      load8(n0 + 8, wb);
      fma8(n0, wa);
      load8(n0 + 16, wa);
      fma8(n0 + 8, wb);
    }
    if (extra) {
      const float w = wextra[t];
#pragma unroll
      for (int r = 0; r < TM; ++r) acc[r] = fmaf(extra[r], w, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < TM; ++r) gt[t][r] = (xmask == nullptr || xmask[r * ldm + t] > 0.f) ? acc[r] : 0.f;
  }
}

template <int TM>
__global__ void __launch_bounds__(SM_THREADS)
small_mlp_bwd_kernel(MlpLayout L, const float* __restrict__ params, const float* __restrict__ g_out, SmallStash s,
                     float* __restrict__ g_params, int64_t M, int S) {
  __shared__ __align__(16) float dza[128][TM];
  __shared__ __align__(16) float dzb[128][TM];
  __shared__ __align__(16) float xin[TM][SM_XLD];
  __shared__ float gsig[TM];
  const int t = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * TM;
  const int W = L.W, W2 = L.W / 2, IC = L.in_ch, ICV = L.in_ch_v, OC = L.out_ch;
  auto load_x = [&](const float* src, int ld, int K, int col0, int row_div) {   // xin[r][col0 + k] = src[row][k]
    for (int idx = t; idx < TM * K; idx += SM_THREADS) {
      const int r = idx / K, k = idx - r * K;
      xin[r][col0 + k] = (m0 + r < M) ? src[((m0 + r) / row_div) * ld + k] : 0.f;
    }
  };
  float (*dz)[TM] = dza;
  float (*gn)[TM] = dzb;
  // ---- rgb head: dZ = g_out[:, 0:3] (rows beyond M are zero), X = hv
  for (int idx = t; idx < 128 * TM; idx += SM_THREADS) (&dza[0][0])[idx] = 0.f;
  __syncthreads();
  for (int idx = t; idx < TM * OC; idx += SM_THREADS) {
    const int r = idx / OC, c = idx - r * OC;
    const float v = (m0 + r < M) ? g_out[(m0 + r) * OC + c] : 0.f;
    if (c < 3) dz[c][r] = v; else gsig[r] = v;
  }
  load_x(s.hv, W2, W2, 0, 1);
  __syncthreads();
  small_wgrad<TM>(t, 3, dz, &xin[0][0], SM_XLD, W2, g_params + L.rgb_w, W2, 0);
  small_bgrad<TM>(t, 3, dz, g_params + L.rgb_b);
  small_dgrad<TM>(t, 3, W2, dz, params + L.rgb_w, W2, 0, &xin[0][0], SM_XLD, nullptr, nullptr, gn);
  __syncthreads();
  { float (*tmp)[TM] = dz; dz = gn; gn = tmp; }
  // ---- views layer: hv = relu([feat | enc_views] Wv^T + b)
  load_x(s.feat, W, W, 0, 1);
  load_x(s.enc_views, ICV, ICV, W, S);
  __syncthreads();
  small_wgrad<TM>(t, W2, dz, &xin[0][0], SM_XLD, W + ICV, g_params + L.views_w, W + ICV, 0);
  small_bgrad<TM>(t, W2, dz, g_params + L.views_b);
  small_dgrad<TM>(t, W2, W, dz, params + L.views_w, W + ICV, 0, nullptr, 0, nullptr, nullptr, gn);   // feature: no activation
  __syncthreads();
  { float (*tmp)[TM] = dz; dz = gn; gn = tmp; }
  // ---- feature layer (+ sigma head): X = h_{D-1}
  load_x(s.h[L.D - 1], W, W, 0, 1);
  __syncthreads();
  small_wgrad<TM>(t, W, dz, &xin[0][0], SM_XLD, W, g_params + L.feat_w, W, 0);
  small_bgrad<TM>(t, W, dz, g_params + L.feat_b);
  const bool has_alpha = (L.out_mode == 0);
  if (has_alpha) {
    if (t < W) {     // dW_alpha[k] += sum_r g_sigma[r] h[r][k]
      float sacc = 0.f;
      for (int r = 0; r < TM; ++r) sacc = fmaf(gsig[r], xin[r][t], sacc);
      atomicAdd(g_params + L.alpha_w + t, sacc);
    }
    if (t == 0) { float sacc = 0.f; for (int r = 0; r < TM; ++r) sacc += gsig[r]; atomicAdd(g_params + L.alpha_b, sacc); }
  }
  small_dgrad<TM>(t, W, W, dz, params + L.feat_w, W, 0, &xin[0][0], SM_XLD, has_alpha ? gsig : nullptr,
                  params + L.alpha_w, gn);
  __syncthreads();
  { float (*tmp)[TM] = dz; dz = gn; gn = tmp; }
  // ---- trunk, last to first: dz = dL/dz_i (ReLU already applied)
  for (int i = L.D - 1; i >= 0; --i) {
    const bool skip = (L.skip >= 0 && i == L.skip + 1);
    const int fan = L.pts_in[i];
    int hcol = 0;                                   // column of h_{i-1} inside the layer input
    if (i == 0) load_x(s.enc_pts, IC, IC, 0, 1);
    else if (skip) { load_x(s.enc_pts, IC, IC, 0, 1); load_x(s.h[i - 1], W, W, IC, 1); hcol = IC; }
    else load_x(s.h[i - 1], W, W, 0, 1);
    __syncthreads();
    small_wgrad<TM>(t, W, dz, &xin[0][0], SM_XLD, fan, g_params + L.pts_w[i], fan, 0);
    small_bgrad<TM>(t, W, dz, g_params + L.pts_b[i]);
    if (i > 0) small_dgrad<TM>(t, W, W, dz, params + L.pts_w[i], fan, hcol, &xin[0][hcol], SM_XLD, nullptr, nullptr, gn);
    __syncthreads();
    { float (*tmp)[TM] = dz; dz = gn; gn = tmp; }
  }
}

}  // namespace lush

using namespace lush;

// the fused path covers: width <= 128 (one thread per unit), in_ch <= 63, few rows (latency-bound regime), and in
// the backward no d/d inputs (the noise MLP's points are detached, models/lushnerf.py:422,614)
bool lush_mlp_small_supported(const MlpLayout& L, int64_t M) {
  return L.W <= 128 && (L.W % 8) == 0 && L.in_ch <= 63 && L.in_ch_v <= 31 && L.W + L.in_ch <= SM_XLD &&
         M <= 32768;
}

int lush_mlp_small_fwd(const MlpLayout& L, const lush_mlp_desc* desc, const float* params, const float* ray_batch,
                       int ray_stride, const float* z, float* out, float* const* stash_ptrs, int64_t M, int S,
                       cudaStream_t st) {
  SmallStash s;
  s.enc_pts = stash_ptrs[0]; s.enc_views = stash_ptrs[1]; s.feat = stash_ptrs[2]; s.hv = stash_ptrs[3];
  for (int i = 0; i < kMaxDepth; ++i) s.h[i] = stash_ptrs[4 + i];
  if (M >= SM_BIG_ROWS)
    small_mlp_fwd_kernel<SM_FWD_TM_BIG><<<(unsigned)ceil_div(M, SM_FWD_TM_BIG), SM_THREADS, 0, st>>>(
        L, params, ray_batch, ray_stride, z, out, s, M, S, desc->L_pts, desc->L_views);
  else
    small_mlp_fwd_kernel<SM_FWD_TM><<<(unsigned)ceil_div(M, SM_FWD_TM), SM_THREADS, 0, st>>>(
        L, params, ray_batch, ray_stride, z, out, s, M, S, desc->L_pts, desc->L_views);
  LUSH_LAUNCH_CHECK("small_mlp_fwd");
  return 0;
}

int lush_mlp_small_bwd(const MlpLayout& L, const float* params, const float* g_out, float* const* stash_ptrs,
                       float* g_params, int64_t M, int S, cudaStream_t st) {
  SmallStash s;
  s.enc_pts = stash_ptrs[0]; s.enc_views = stash_ptrs[1]; s.feat = stash_ptrs[2]; s.hv = stash_ptrs[3];
  for (int i = 0; i < kMaxDepth; ++i) s.h[i] = stash_ptrs[4 + i];
  small_mlp_bwd_kernel<SM_BWD_TM><<<(unsigned)ceil_div(M, SM_BWD_TM), SM_THREADS, 0, st>>>(L, params, g_out, s, g_params,
                                                                                         M, S);
  LUSH_LAUNCH_CHECK("small_mlp_bwd");
  return 0;
}
