// Exact-fp32 SIMT implementation of the positional encoder + NeRF / NeRF_Noise MLPs, forward and
// backward (impl = 0 of lush_mlp_fwd / lush_mlp_bwd).  This is the strict-parity mode: every product
// is an fp32 FFMA with fp32 accumulation, matching the reference's "true fp32" cuBLAS SGEMM path
// (SURVEY.md 2.2).  It also serves networks whose width the tcgen05 path does not cover (the 4x128
// noise MLP).  Replaces reference helpers:311-344, 394-423, 483-512 and models/lushnerf.py:234-293.
//
// Structure: one generic 64x64x16 register-tiled GEMM kernel with three operand-major modes
//   FWD   C[m,n] = act( sum_k [X1|X2][m,k] W[n,k] + b[n] )
//   DGRAD C[m,k] = ( sum_n G[m,n] W[n,k] ) * (Mask[m,k] > 0)        (optionally accumulating)
//   WGRAD C[n,k] += sum_m G[m,n] X[m,k]                              (split over m, fp32 atomics)
// plus encode / encode-backward / bias-grad / relu-mask kernels.
#include <stdlib.h>
#include "mlp_common.cuh"

namespace lush {

constexpr int BM = 64, BN = 64, BK = 16;

struct Seg {            // one K-segment of the "row" operand (concatenated inputs)
  const float* ptr;
  int ld;               // row stride (floats)
  int K;                // width of this segment
  int row_div;          // physical row = logical row / row_div (per-ray tensors broadcast to samples)
};

struct GemmArgs {
  Seg a1, a2;           // A(i, kk): kk < a1.K -> a1 else a2   (A_RED_CONTIG)  |  a1 only when !A_RED_CONTIG
  const float* B; int ldb; int b_row_div;   // B operand
  float* C; int ldc;
  int64_t Mout; int Nout; int64_t Kred;
  const float* bias;    // per output column (FWD), nullable
  int relu;
  const float* mask; int ldmask;   // multiply output by (mask[i,j] > 0), nullable
  int beta;             // 1: C += result (non-atomic), 0: overwrite
  int atomic;           // 1: atomicAdd result into C (split-K)
  int64_t kchunk;       // reduction elements per blockIdx.z
};

// A(i,kk): A_RED_CONTIG -> seg[(i/div)*ld + kk];  else (WGRAD) a1.ptr[kk*ld + i]
// B(kk,j): B_RED_CONTIG -> B[j*ldb + kk];         else B[(kk/b_row_div)*ldb + j]
template <bool A_RED_CONTIG, bool B_RED_CONTIG>
__global__ void __launch_bounds__(256) gemm_kernel(GemmArgs g) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * g.kchunk;
  const int64_t kend = min(g.Kred, kbeg + g.kchunk);
  const int tm = tid / 16, tn = tid % 16;  // 16x16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- load A tile (BM x BK) into As[k][m]
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      int mi, ki;
      if (A_RED_CONTIG) { ki = tid % 16; mi = tid / 16 + 16 * it; }
      else { mi = tid % 64; ki = tid / 64 + 4 * it; }
      int64_t m = m0 + mi, kk = k0 + ki;
      float v = 0.f;
      if (m < g.Mout && kk < kend) {
        if (A_RED_CONTIG) {
          if (kk < g.a1.K) v = g.a1.ptr[(m / g.a1.row_div) * g.a1.ld + kk];
          else v = g.a2.ptr[(m / g.a2.row_div) * g.a2.ld + (kk - g.a1.K)];
        } else {
          v = g.a1.ptr[kk * g.a1.ld + m];
        }
      }
      As[ki][mi] = v;
    }
    // ---- load B tile (BK x BN) into Bs[k][n]
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      int ni, ki;
      if (B_RED_CONTIG) { ki = tid % 16; ni = tid / 16 + 16 * it; }
      else { ni = tid % 64; ki = tid / 64 + 4 * it; }
      int n = n0 + ni;
      int64_t kk = k0 + ki;
      float v = 0.f;
      if (n < g.Nout && kk < kend) {
        if (B_RED_CONTIG) v = g.B[(int64_t)n * g.ldb + kk];
        else v = g.B[(kk / g.b_row_div) * g.ldb + n];
      }
      Bs[ki][ni] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][tm * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + tm * 4 + i;
    if (m >= g.Mout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tn * 4 + j;
      if (n >= g.Nout) continue;
      float v = acc[i][j];
      float* c = g.C + m * g.ldc + n;
      if (g.atomic) { atomicAdd(c, v); continue; }
      if (g.bias) v += g.bias[n];
      if (g.beta) v += *c;
      if (g.relu) v = fmaxf(v, 0.f);
      if (g.mask) v = (g.mask[m * g.ldmask + n] > 0.f) ? v : 0.f;
      *c = v;
    }
  }
}

// ---------------------------------------------------------------- encoder
// enc_pts [M, in_ch]: [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)] (helpers:334-344)
__global__ void encode_pts_kernel(const float* __restrict__ ray_batch, int ray_stride,
                                  const float* __restrict__ z, float* __restrict__ enc, int64_t m_beg,
                                  int64_t m_cnt, int S, int L) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= m_cnt * 3) return;
  int64_t ml = i / 3;
  int c = (int)(i - ml * 3);
  int64_t m = m_beg + ml, r = m / S;
  const float* rb = ray_batch + r * ray_stride;
  float x = __fadd_rn(rb[c], __fmul_rn(rb[3 + c], z[m]));  // o + d*z, unfused like the reference
  int in_ch = 3 + 6 * L;
  float* e = enc + ml * in_ch;
  e[c] = x;
  float f = 1.f;
  for (int l = 0; l < L; ++l) {
    float s, cs;
    sincosf(x * f, &s, &cs);
    e[3 + 6 * l + c] = s;
    e[3 + 6 * l + 3 + c] = cs;
    f *= 2.f;
  }
}

__global__ void encode_views_kernel(const float* __restrict__ ray_batch, int ray_stride,
                                    float* __restrict__ enc, int64_t R, int L) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= R * 3) return;
  int64_t r = i / 3;
  int c = (int)(i - r * 3);
  float x = ray_batch[r * ray_stride + 8 + c];
  int in_ch = 3 + 6 * L;
  float* e = enc + r * in_ch;
  e[c] = x;
  float f = 1.f;
  for (int l = 0; l < L; ++l) {
    float s, cs;
    sincosf(x * f, &s, &cs);
    e[3 + 6 * l + c] = s;
    e[3 + 6 * l + 3 + c] = cs;
    f *= 2.f;
  }
}

// d enc -> d x for one scalar coordinate
__device__ __forceinline__ float encode_bwd_scalar(float x, const float* __restrict__ ge, int c, int L) {
  float gx = ge[c];
  float f = 1.f;
  for (int l = 0; l < L; ++l) {
    float s, cs;
    sincosf(x * f, &s, &cs);
    gx += f * (cs * ge[3 + 6 * l + c] - s * ge[3 + 6 * l + 3 + c]);
    f *= 2.f;
  }
  return gx;
}

// one warp per ray: reduces d enc_pts over the ray's samples into d o, d d and the per-sample
// d enc_views [M, in_ch_v] into d viewdirs.  g_ray_batch [R,11].
__global__ void encode_bwd_kernel(const float* __restrict__ ray_batch, int ray_stride,
                                  const float* __restrict__ z, const float* __restrict__ g_enc_pts,
                                  const float* __restrict__ g_enc_views, float* __restrict__ g_rb,
                                  int64_t R, int S, int L, int Lv) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const float* rb = ray_batch + r * ray_stride;
  const int in_ch = 3 + 6 * L, in_ch_v = 3 + 6 * Lv;
  float go[3] = {0, 0, 0}, gd[3] = {0, 0, 0}, gv[3] = {0, 0, 0};
  for (int s = lane; s < S; s += 32) {
    int64_t m = r * S + s;
    float zz = z[m];
    for (int c = 0; c < 3; ++c) {
      float x = __fadd_rn(rb[c], __fmul_rn(rb[3 + c], zz));
      float gx = encode_bwd_scalar(x, g_enc_pts + m * in_ch, c, L);
      go[c] += gx;
      gd[c] += gx * zz;
    }
    if (g_enc_views) {
      // every sample of the ray saw the same encoded direction: sum the per-sample grads first
      for (int c = 0; c < 3; ++c)
        gv[c] += encode_bwd_scalar(rb[8 + c], g_enc_views + m * in_ch_v, c, Lv);
    }
  }
  for (int c = 0; c < 3; ++c) { go[c] = warp_sum(go[c]); gd[c] = warp_sum(gd[c]); gv[c] = warp_sum(gv[c]); }
  if (lane == 0) {
    float* o = g_rb + r * 11;
    o[0] = go[0]; o[1] = go[1]; o[2] = go[2]; o[3] = gd[0]; o[4] = gd[1]; o[5] = gd[2];
    o[6] = 0.f; o[7] = 0.f; o[8] = gv[0]; o[9] = gv[1]; o[10] = gv[2];
  }
}

// db[n] += sum_m G[m,n]
__global__ void bias_grad_kernel(const float* __restrict__ G, int ldg, int64_t M, int N,
                                 float* __restrict__ db, int rows_per_block) {
  int n = blockIdx.y * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int64_t mb = (int64_t)blockIdx.x * rows_per_block, me = min(M, mb + rows_per_block);
  float s = 0.f;
  for (int64_t m = mb; m < me; ++m) s += G[m * ldg + n];
  atomicAdd(db + n, s);
}

// G[m,n] = (H[m,n] > 0) ? G[m,n] : 0
__global__ void relu_mask_kernel(float* __restrict__ G, const float* __restrict__ H, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) G[i] = (H[i] > 0.f) ? G[i] : 0.f;
}

// ---------------------------------------------------------------- host-side helpers
struct Launcher {
  cudaStream_t st;
  int rc = 0;
  void fwd(Seg a1, Seg a2, const float* Wt, int ldw, const float* bias, int relu, float* C, int ldc,
           int64_t M, int N) {
    if (rc) return;
    GemmArgs g{};
    g.a1 = a1; g.a2 = a2; g.B = Wt; g.ldb = ldw; g.b_row_div = 1; g.C = C; g.ldc = ldc;
    g.Mout = M; g.Nout = N; g.Kred = a1.K + a2.K; g.bias = bias; g.relu = relu;
    g.kchunk = g.Kred;
    dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(N, BN), 1);
    gemm_kernel<true, true><<<grid, 256, 0, st>>>(g);
    count_launch();
    rc = check_cuda(cudaGetLastError(), "mlp_simt gemm fwd");
  }
  // C[m, 0:Kout] (=|+=) G[m, 0:N] . W[0:N, koff:koff+Kout], optional relu-mask by `mask`
  void dgrad(const float* G, int ldg, int N, const float* Wt, int ldw, float* C, int ldc, int64_t M,
             int Kout, int beta, const float* mask, int ldmask) {
    if (rc) return;
    GemmArgs g{};
    g.a1 = Seg{G, ldg, N, 1}; g.a2 = Seg{nullptr, 0, 0, 1};
    g.B = Wt; g.ldb = ldw; g.b_row_div = 1; g.C = C; g.ldc = ldc;
    g.Mout = M; g.Nout = Kout; g.Kred = N; g.beta = beta; g.mask = mask; g.ldmask = ldmask;
    g.kchunk = g.Kred;
    dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(Kout, BN), 1);
    gemm_kernel<true, false><<<grid, 256, 0, st>>>(g);
    count_launch();
    rc = check_cuda(cudaGetLastError(), "mlp_simt gemm dgrad");
  }
  // dW[0:N, 0:K] += G[0:M, 0:N]^T . X[0:M(/div), 0:K]   (dW row stride lddw)
  void wgrad(const float* G, int ldg, int N, const float* X, int ldx, int x_row_div, int K, float* dW,
             int lddw, int64_t M) {
    if (rc) return;
    GemmArgs g{};
    g.a1 = Seg{G, ldg, 0, 1}; g.a2 = Seg{nullptr, 0, 0, 1};
    g.B = X; g.ldb = ldx; g.b_row_div = x_row_div; g.C = dW; g.ldc = lddw;
    g.Mout = N; g.Nout = K; g.Kred = M; g.atomic = 1;
    // enough (output tile, sample chunk) blocks to fill the 148 SMs twice, also when M is small
    const int64_t bxy = ceil_div(N, BM) * ceil_div(K, BN);
    int64_t chunk = ceil_div(M, ceil_div(2 * kNumSMs, bxy));
    chunk = ((chunk + 15) / 16) * 16;
    if (chunk < 64) chunk = 64;
    if (chunk > 4096) chunk = 4096;
    g.kchunk = chunk;
    dim3 grid((unsigned)ceil_div(N, BM), (unsigned)ceil_div(K, BN), (unsigned)ceil_div(M, chunk));
    gemm_kernel<false, false><<<grid, 256, 0, st>>>(g);
    count_launch();
    rc = check_cuda(cudaGetLastError(), "mlp_simt gemm wgrad");
  }
  void bias_grad(const float* G, int ldg, int64_t M, int N, float* db) {
    if (rc) return;
    int rows = (int)ceil_div(M, 2 * kNumSMs);
    if (rows < 16) rows = 16;
    if (rows > 2048) rows = 2048;
    dim3 grid((unsigned)ceil_div(M, rows), (unsigned)ceil_div(N, 128));
    bias_grad_kernel<<<grid, 128, 0, st>>>(G, ldg, M, N, db, rows);
    count_launch();
    rc = check_cuda(cudaGetLastError(), "mlp_simt bias_grad");
  }
};

// stash layout for rows [0,M): enc_pts [M,in_ch] | enc_views [R,in_ch_v] | h_0..h_{D-1} [M,W] each |
// feat [M,W] | hv [M,W/2]
struct SimtStash {
  float *enc_pts, *enc_views, *h[kMaxDepth], *feat, *hv;
  size_t floats;
};
static SimtStash carve_stash(const MlpLayout& L, float* base, int64_t M, int64_t R) {
  SimtStash s;
  size_t off = 0;
  auto take = [&](size_t n) { float* p = base ? base + off : nullptr; off += (n + 3) & ~(size_t)3; return p; };
  s.enc_pts = take((size_t)M * L.in_ch);
  s.enc_views = take((size_t)R * L.in_ch_v);
  for (int i = 0; i < L.D; ++i) s.h[i] = take((size_t)M * L.W);
  s.feat = take((size_t)M * L.W);
  s.hv = take((size_t)M * (L.W / 2));
  s.floats = off;
  return s;
}

constexpr int64_t kInferChunkRows = 1 << 16;  // same row chunk as the reference netchunk (lushnerf.py:234)

}  // namespace lush

using namespace lush;

// fused small-network path (mlp_small.cu)
bool lush_mlp_small_supported(const MlpLayout& L, int64_t M);
int lush_mlp_small_fwd(const MlpLayout& L, const lush_mlp_desc* desc, const float* params, const float* ray_batch,
                       int ray_stride, const float* z, float* out, float* const* stash_ptrs, int64_t M, int S,
                       cudaStream_t st);
int lush_mlp_small_bwd(const MlpLayout& L, const float* params, const float* g_out, float* const* stash_ptrs,
                       float* g_params, int64_t M, int S, cudaStream_t st);
static void small_stash_ptrs(const SimtStash* s, float** p) {
  for (int i = 0; i < 4 + kMaxDepth; ++i) p[i] = nullptr;
  if (s == nullptr) return;
  p[0] = s->enc_pts; p[1] = s->enc_views; p[2] = s->feat; p[3] = s->hv;
  for (int i = 0; i < kMaxDepth; ++i) p[4 + i] = s->h[i];
}
static bool small_path_enabled() {
  const char* e = getenv("LUSH_SMALL_MLP");   // LUSH_SMALL_MLP=0: always use the layer-by-layer GEMM path
  return !(e && e[0] == '0');
}

size_t lush_mlp_simt_stash_bytes(const MlpLayout& L, int64_t R, int S) {
  return carve_stash(L, nullptr, R * S, R).floats * sizeof(float);
}
size_t lush_mlp_simt_workspace_bytes(const MlpLayout& L, int64_t R, int S, int training) {
  int64_t M = R * S;
  if (training) {
    // backward scratch: two gradient ping-pong buffers [M, W + in_ch] + per-sample view-enc grads
    return ((size_t)M * (2 * (size_t)(L.W + L.in_ch) + L.in_ch_v) + 16) * sizeof(float);
  }
  int64_t rows = M < kInferChunkRows ? M : kInferChunkRows;
  return carve_stash(L, nullptr, rows, R).floats * sizeof(float);
}

// forward over rows [m_beg, m_beg+m_cnt) with activations written into `s` (row-local indexing)
static int simt_forward_rows(const MlpLayout& L, const float* params, const float* ray_batch,
                             int ray_stride, const float* z, float* out, const SimtStash& s,
                             int64_t m_beg, int64_t m_cnt, int S, int Lp, Launcher& go) {
  const int W = L.W, W2 = L.W / 2;
  {
    int th = 256;
    encode_pts_kernel<<<(unsigned)ceil_div(m_cnt * 3, th), th, 0, go.st>>>(ray_batch, ray_stride, z,
                                                                          s.enc_pts, m_beg, m_cnt, S, Lp);
    count_launch();
    int rc = check_cuda(cudaGetLastError(), "encode_pts");
    if (rc) return rc;
  }
  Seg none{nullptr, 0, 0, 1};
  const float* h_prev = nullptr;
  for (int i = 0; i < L.D; ++i) {
    Seg a1, a2 = none;
    if (i == 0) a1 = Seg{s.enc_pts, L.in_ch, L.in_ch, 1};
    else if (L.skip >= 0 && i == L.skip + 1) { a1 = Seg{s.enc_pts, L.in_ch, L.in_ch, 1}; a2 = Seg{h_prev, W, W, 1}; }
    else a1 = Seg{h_prev, W, W, 1};
    go.fwd(a1, a2, params + L.pts_w[i], L.pts_in[i], params + L.pts_b[i], 1, s.h[i], W, m_cnt, W);
    h_prev = s.h[i];
  }
  // feature (no activation), views layer on [feature | enc_views(ray)], rgb head; alpha head
  go.fwd(Seg{h_prev, W, W, 1}, none, params + L.feat_w, W, params + L.feat_b, 0, s.feat, W, m_cnt, W);
  // enc_views is indexed by GLOBAL ray = (m_beg + ml)/S; chunks start at multiples of S only when
  // m_beg % S == 0, so pass a pointer pre-offset to the first ray of the chunk and require alignment.
  go.fwd(Seg{s.feat, W, W, 1}, Seg{s.enc_views + (m_beg / S) * L.in_ch_v, L.in_ch_v, L.in_ch_v, S},
         params + L.views_w, W + L.in_ch_v, params + L.views_b, 1, s.hv, W2, m_cnt, W2);
  float* o = out + m_beg * L.out_ch;
  go.fwd(Seg{s.hv, W2, W2, 1}, none, params + L.rgb_w, W2, params + L.rgb_b, 0, o, L.out_ch, m_cnt, 3);
  if (L.out_mode == 0)
    go.fwd(Seg{h_prev, W, W, 1}, none, params + L.alpha_w, W, params + L.alpha_b, 0, o + 3, L.out_ch,
           m_cnt, 1);
  return go.rc;
}

int lush_mlp_simt_fwd(const MlpLayout& L, const lush_mlp_desc* desc, const float* params,
                      const float* ray_batch, int ray_stride, const float* z, float* out, void* stash,
                      void* workspace, size_t workspace_bytes, int64_t R, int S, cudaStream_t st) {
  const int64_t M = R * S;
  Launcher go{st};
  if (small_path_enabled() && lush_mlp_small_supported(L, M)) {   // few rows of a narrow net: one fused launch
    float* p[4 + kMaxDepth];
    if (stash != nullptr) {
      SimtStash s = carve_stash(L, (float*)stash, M, R);
      for (int i = L.D; i < kMaxDepth; ++i) s.h[i] = nullptr;
      small_stash_ptrs(&s, p);
    } else {
      small_stash_ptrs(nullptr, p);
    }
    return lush_mlp_small_fwd(L, desc, params, ray_batch, ray_stride, z, out, p, M, S, st);
  }
  // encoded view directions, once per ray (the reference re-encodes them per sample, lushnerf.py:241-243)
  if (stash != nullptr) {
    SimtStash s = carve_stash(L, (float*)stash, M, R);
    int th = 256;
    encode_views_kernel<<<(unsigned)ceil_div(R * 3, th), th, 0, st>>>(ray_batch, ray_stride, s.enc_views, R,
                                                                     desc->L_views);
    count_launch();
    int rc = check_cuda(cudaGetLastError(), "encode_views");
    if (rc) return rc;
    return simt_forward_rows(L, params, ray_batch, ray_stride, z, out, s, 0, M, S, desc->L_pts, go);
  }
  // inference: chunk rows (multiple of S) through a small workspace
  int64_t rows = M < kInferChunkRows ? M : kInferChunkRows;
  rows = (rows / S) * S;
  if (rows == 0) rows = S;
  size_t need = carve_stash(L, nullptr, rows, R).floats * sizeof(float);
  if (workspace == nullptr || workspace_bytes < need) {
    // rows was rounded DOWN to a multiple of S so need <= the queried size unless S > chunk
    if (workspace == nullptr || workspace_bytes < need) {
      set_error("mlp_fwd(simt): workspace too small (%zu < %zu)", workspace_bytes, need);
      return LUSH_EWORKSPACE;
    }
  }
  SimtStash s = carve_stash(L, (float*)workspace, rows, R);
  {
    int th = 256;
    encode_views_kernel<<<(unsigned)ceil_div(R * 3, th), th, 0, st>>>(ray_batch, ray_stride, s.enc_views, R,
                                                                     desc->L_views);
    count_launch();
    int rc = check_cuda(cudaGetLastError(), "encode_views");
    if (rc) return rc;
  }
  for (int64_t m = 0; m < M; m += rows) {
    int64_t cnt = (M - m) < rows ? (M - m) : rows;
    int rc = simt_forward_rows(L, params, ray_batch, ray_stride, z, out, s, m, cnt, S, desc->L_pts, go);
    if (rc) return rc;
  }
  return 0;
}

int lush_mlp_simt_bwd(const MlpLayout& L, const lush_mlp_desc* desc, const float* params,
                      const float* ray_batch, int ray_stride, const float* z, const float* g_out,
                      const void* stash, float* g_params, float* g_ray_batch, void* workspace,
                      size_t workspace_bytes, int64_t R, int S, cudaStream_t st) {
  const int64_t M = R * S;
  const int W = L.W, W2 = L.W / 2, IC = L.in_ch, ICV = L.in_ch_v;
  size_t need = lush_mlp_simt_workspace_bytes(L, R, S, 1);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("mlp_bwd(simt): workspace too small (%zu < %zu)", workspace_bytes, need);
    return LUSH_EWORKSPACE;
  }
  SimtStash s = carve_stash(L, (float*)const_cast<void*>(stash), M, R);
  if (g_ray_batch == nullptr && small_path_enabled() && lush_mlp_small_supported(L, M)) {
    float* p[4 + kMaxDepth];
    for (int i = L.D; i < kMaxDepth; ++i) s.h[i] = nullptr;
    small_stash_ptrs(&s, p);
    return lush_mlp_small_bwd(L, params, g_out, p, g_params, M, S, st);
  }
  float* gA = (float*)workspace;                 // [M, W]   gradient wrt current activation
  float* gB = gA + (size_t)M * W;                // [M, W]
  float* g_enc = gB + (size_t)M * W;             // [M, in_ch]   (two contributions: layer 0, skip layer)
  float* g_encv = g_enc + (size_t)M * IC * 2;    // [M, in_ch_v] per-sample view-enc grads
  // layout note: 2*(W+IC) floats per row were reserved: gA, gB (2W) + g_enc (IC) + spare IC
  Launcher go{st};
  const bool need_in = g_ray_batch != nullptr;
  const int OC = L.out_ch;
  const float* h_last = s.h[L.D - 1];

  // ---- rgb head: out[:,0:3] = hv . Wrgb^T + b
  go.wgrad(g_out, OC, 3, s.hv, W2, 1, W2, g_params + L.rgb_w, W2, M);
  go.bias_grad(g_out, OC, M, 3, g_params + L.rgb_b);
  // g_hv (masked by hv > 0) -> gA[:, 0:W2]
  go.dgrad(g_out, OC, 3, params + L.rgb_w, W2, gA, W2, M, W2, 0, s.hv, W2);
  // ---- views layer: hv = relu([feat | enc_views] . Wv^T + b)
  go.wgrad(gA, W2, W2, s.feat, W, 1, W, g_params + L.views_w, W + ICV, M);
  go.wgrad(gA, W2, W2, s.enc_views, ICV, S, ICV, g_params + L.views_w + W, W + ICV, M);
  go.bias_grad(gA, W2, M, W2, g_params + L.views_b);
  if (need_in) go.dgrad(gA, W2, W2, params + L.views_w + W, W + ICV, g_encv, ICV, M, ICV, 0, nullptr, 0);
  // g_feat -> gB [M, W] (feature_linear has no activation)
  go.dgrad(gA, W2, W2, params + L.views_w, W + ICV, gB, W, M, W, 0, nullptr, 0);
  // ---- feature layer: feat = h_last . Wf^T + b
  go.wgrad(gB, W, W, h_last, W, 1, W, g_params + L.feat_w, W, M);
  go.bias_grad(gB, W, M, W, g_params + L.feat_b);
  // g_h_last = gB . Wf  (+ alpha head contribution), then relu mask of the last trunk layer
  const bool has_alpha = (L.out_mode == 0);
  go.dgrad(gB, W, W, params + L.feat_w, W, gA, W, M, W, 0, has_alpha ? nullptr : h_last, W);
  if (has_alpha) {
    go.wgrad(g_out + 3, OC, 1, h_last, W, 1, W, g_params + L.alpha_w, W, M);
    go.bias_grad(g_out + 3, OC, M, 1, g_params + L.alpha_b);
    go.dgrad(g_out + 3, OC, 1, params + L.alpha_w, W, gA, W, M, W, 1, h_last, W);
  }
  // ---- trunk, last to first.  gcur holds dL/dz_i (already relu-masked)
  float* gcur = gA;
  float* gnext = gB;
  bool enc_has = false;
  for (int i = L.D - 1; i >= 0; --i) {
    const float* Wi = params + L.pts_w[i];
    float* dWi = g_params + L.pts_w[i];
    const int fan = L.pts_in[i];
    go.bias_grad(gcur, W, M, W, g_params + L.pts_b[i]);
    if (i == 0) {
      go.wgrad(gcur, W, W, s.enc_pts, IC, 1, IC, dWi, fan, M);
      if (need_in) go.dgrad(gcur, W, W, Wi, fan, g_enc, IC, M, IC, enc_has ? 1 : 0, nullptr, 0);
    } else if (L.skip >= 0 && i == L.skip + 1) {
      // input = [enc_pts | h_{i-1}]
      go.wgrad(gcur, W, W, s.enc_pts, IC, 1, IC, dWi, fan, M);
      go.wgrad(gcur, W, W, s.h[i - 1], W, 1, W, dWi + IC, fan, M);
      if (need_in) { go.dgrad(gcur, W, W, Wi, fan, g_enc, IC, M, IC, 0, nullptr, 0); enc_has = true; }
      go.dgrad(gcur, W, W, Wi + IC, fan, gnext, W, M, W, 0, s.h[i - 1], W);
      float* t = gcur; gcur = gnext; gnext = t;
    } else {
      go.wgrad(gcur, W, W, s.h[i - 1], W, 1, W, dWi, fan, M);
      go.dgrad(gcur, W, W, Wi, fan, gnext, W, M, W, 0, s.h[i - 1], W);
      float* t = gcur; gcur = gnext; gnext = t;
    }
  }
  if (go.rc) return go.rc;
  if (need_in) {
    int warps = 4;
    encode_bwd_kernel<<<(unsigned)ceil_div(R, warps), warps * 32, 0, st>>>(
        ray_batch, ray_stride, z, g_enc, g_encv, g_ray_batch, R, S, desc->L_pts, desc->L_views);
    count_launch();
    int rc = check_cuda(cudaGetLastError(), "encode_bwd");
    if (rc) return rc;
  }
  return 0;
}
