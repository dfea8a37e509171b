// Motion-parameter network of the rigid blurring kernel, forward and backward in ONE launch each.
// Replaces Rigid_Blurring_Kernel.forward up to the warp (reference models/lushnerf.py:118-147): view embedding ->
// trunk (ReLU, optional skip concat) -> r / v / w branches -> r*window, v*window, w = sigmoid, w / (sum w + 1e-10).
//
// Everything up to (r, v, ccw) depends on the image index alone, so the network runs once per IMAGE (rows of
// the embedding table) and every ray gathers the row of its image; the backward sums the per-ray gradients per
// image first.  A CTA owns MOTION_ROWS images and keeps every activation of those rows in shared memory through
// the whole network, next to a copy of all weights and biases (~100 KB, staged once per CTA with every load in
// flight); a warp owns an output unit, its lanes stride the (contiguous) fan-in of that unit.  ~25 k parameters:
// the kernels are latency bound by design -- they replace ~90 library launches of a few microseconds each.
#include "common.cuh"

namespace lush {

constexpr int MOTION_ROWS = 8;
constexpr int MOTION_THREADS = 512;
constexpr int MOTION_CHUNK = 4096;       // rays whose image slot is staged in shared memory at a time

// The layer table sits in kernel-parameter (constant) space and is indexed with a runtime layer number: every field
// access is an indexed constant load with ~50 cycles of latency.  Each layer's scalars are therefore read ONCE into
// registers (LayerRegs) before its loops.
struct LayerRegs {
  const float* w; const float* b; float* gw; float* gb;
  int in0_off, in0_dim, in1_off, in1_dim, out_off, out_dim, act;
};
__device__ __forceinline__ LayerRegs motion_layer(const lush_motion_net& net, int l) {
  const lush_dense_layer& L = net.layers[l];
  LayerRegs r;
  r.w = L.w; r.b = L.b; r.gw = L.gw; r.gb = L.gb;
  r.in0_off = L.in0_off; r.in0_dim = L.in0_dim; r.in1_off = L.in1_off; r.in1_dim = L.in1_dim;
  r.out_off = L.out_off; r.out_dim = L.out_dim; r.act = L.act;
  return r;
}
__device__ __forceinline__ int motion_src(const LayerRegs& L, int k) {
  return k < L.in0_dim ? L.in0_off + k : L.in1_off + (k - L.in0_dim);
}

// Sums of 8 per-lane accumulators (one per row) over the warp in 9 shuffles instead of 40: the lanes trade halves of
// their rows while they reduce (xor 16, 8, 4), then finish the one row they are left with (xor 2, 1).
static_assert(MOTION_ROWS == 8, "motion_reduce_rows is written for 8 rows");
__device__ __forceinline__ int motion_row_of(int lane) { return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); }
__device__ __forceinline__ float motion_reduce_rows(float (&acc)[MOTION_ROWS], int lane) {
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float send = h16 ? acc[r] : acc[r + 4], keep = h16 ? acc[r + 4] : acc[r];
    acc[r] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const float send = h8 ? acc[r] : acc[r + 2], keep = h8 ? acc[r + 2] : acc[r];
    acc[r] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  const float send = h4 ? acc[0] : acc[1], keep = h4 ? acc[1] : acc[0];
  float v = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

// All weight matrices (~100 KB) are copied into shared memory once per CTA (coalesced, every load in flight at the
// same time): the layer loops then read them with shared-memory latency instead of one L2 round trip per output unit.
// `stage_w` = 0 (network too large for shared memory): read them from global memory in place.
__device__ __forceinline__ void motion_stage_weights(const lush_motion_net& net, float* s_w, int tid) {
  int base = 0;
  for (int l = 0; l < net.n_layers; ++l) {
    const lush_dense_layer& L = net.layers[l];
    const int total = L.out_dim * (L.in0_dim + L.in1_dim);
#pragma unroll 8
    for (int i = tid; i < total; i += MOTION_THREADS) s_w[base + i] = L.w[i];
    base += total;
  }
  for (int l = 0; l < net.n_layers; ++l) {               // biases behind the weights
    const lush_dense_layer& L = net.layers[l];
    for (int i = tid; i < L.out_dim; i += MOTION_THREADS) s_w[base + i] = L.b[i];
    base += L.out_dim;
  }
}
__device__ __forceinline__ int motion_bias_offset(const lush_motion_net& net, int l) {
  int base = 0;
  for (int i = 0; i < net.n_layers; ++i) base += net.layers[i].out_dim * (net.layers[i].in0_dim + net.layers[i].in1_dim);
  for (int i = 0; i < l; ++i) base += net.layers[i].out_dim;
  return base;
}
__device__ __forceinline__ int motion_weight_offset(const lush_motion_net& net, int l) {
  int base = 0;
  for (int i = 0; i < l; ++i) base += net.layers[i].out_dim * (net.layers[i].in0_dim + net.layers[i].in1_dim);
  return base;
}

__global__ void __launch_bounds__(MOTION_THREADS)
motion_fwd_kernel(const __grid_constant__ lush_motion_net net, const float* __restrict__ embed, int n_img,
                  const int64_t* __restrict__ idx, int64_t N, float* __restrict__ rows,
                  float* __restrict__ out_r, float* __restrict__ out_v, float* __restrict__ out_ccw, int stage_w,
                  int* __restrict__ err_flag) {
  extern __shared__ float s_act[];                      // [MOTION_ROWS][rw]
  __shared__ int s_slot[MOTION_CHUNK];
  const int rw = net.row_width, C = net.embed_dim, M = net.num_motion;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = MOTION_THREADS / 32;
  const int row0 = blockIdx.x * MOTION_ROWS;
  const int nrows = min(MOTION_ROWS, n_img - row0);
  float* s_w = s_act + MOTION_ROWS * rw;                // staged weights (stage_w)
  if (stage_w) motion_stage_weights(net, s_w, tid);
  for (int t = tid; t < MOTION_ROWS * C; t += MOTION_THREADS) {
    int r = t / C, k = t - r * C;
    s_act[r * rw + k] = r < nrows ? embed[(int64_t)(row0 + r) * C + k] : 0.f;
  }
  __syncthreads();
  int woff = 0, boff = motion_bias_offset(net, 0);       // running offsets into the staged weights / biases
  for (int l = 0; l < net.n_layers; ++l) {
    const LayerRegs L = motion_layer(net, l);
    const int in_dim = L.in0_dim + L.in1_dim;
    for (int o = warp; o < L.out_dim; o += nwarps) {
      float acc[MOTION_ROWS];
#pragma unroll
      for (int r = 0; r < MOTION_ROWS; ++r) acc[r] = 0.f;
      const float* wrow = (stage_w ? s_w + woff : L.w) + (int64_t)o * in_dim;
#pragma unroll 2
      for (int k = lane; k < in_dim; k += 32) {
        float w = wrow[k];
        const float* a = s_act + motion_src(L, k);
#pragma unroll
        for (int r = 0; r < MOTION_ROWS; ++r) acc[r] = fmaf(w, a[r * rw], acc[r]);
      }
      const float mine = motion_reduce_rows(acc, lane) + (stage_w ? s_w[boff + o] : L.b[o]);       // lane holds row motion_row_of(lane)
      if ((lane & 3) == 0) s_act[motion_row_of(lane) * rw + L.out_off + o] = L.act ? fmaxf(mine, 0.f) : mine;
    }
    woff += L.out_dim * in_dim;
    boff += L.out_dim;
    __syncthreads();
  }
  // stash the rows (pre-head values) for the backward
  for (int t = tid; t < nrows * rw; t += MOTION_THREADS) rows[(int64_t)row0 * rw + t] = s_act[t];
  __syncthreads();
  // heads: r, v scaled by the window; w -> sigmoid -> normalised composition weights
  for (int t = tid; t < MOTION_ROWS * 6 * M; t += MOTION_THREADS) {
    int r = t / (6 * M), j = t - r * 6 * M;
    int off = j < 3 * M ? net.r_off + j : net.v_off + (j - 3 * M);
    s_act[r * rw + off] *= net.rv_window;
  }
  if (tid < MOTION_ROWS) {
    float* a = s_act + tid * rw + net.w_off;
    float sum = 0.f;
    for (int j = 0; j <= M; ++j) {
      a[j] = sigmoidf_(a[j]);
      sum += a[j];
    }
    sum += 1e-10f;
    for (int j = 0; j <= M; ++j) a[j] = a[j] / sum;
  }
  __syncthreads();
  // every ray gathers the row of its image (each CTA serves the rays of its own images): the image slots of a chunk
  // of rays are staged in shared memory first, so the copy loop has no dependent global load in it
  const int nr = 3 * M, nc = M + 1;
  for (int64_t n0 = 0; n0 < N; n0 += MOTION_CHUNK) {
    const int cn = (int)min((int64_t)MOTION_CHUNK, N - n0);
    __syncthreads();
    for (int i = tid; i < cn; i += MOTION_THREADS) {
      const int64_t img = idx[n0 + i], r = img - row0;
      int slot = (r >= 0 && r < nrows) ? (int)r : -1;
      // An image index outside [0, n_img) belongs to no CTA (the reference's nn.Embedding raises on it,
      // models/lushnerf.py:32,119).  CTA 0 gives such a ray the identity kernel (r = v = 0, all weight on the original
      // ray) so that its outputs are defined, and raises the caller's error flag.
      if (blockIdx.x == 0 && (img < 0 || img >= n_img)) {
        slot = -2;
        if (err_flag) atomicOr(err_flag, LUSH_FLAG_BAD_IMAGE_INDEX);
      }
      s_slot[i] = slot;
    }
    __syncthreads();
#pragma unroll 4
    for (int t = tid; t < cn * nr; t += MOTION_THREADS) {   // consecutive threads -> consecutive output floats
      const int i = t / nr, j = t - i * nr;
      const int r = s_slot[i];
      if (r >= 0) {
        out_r[n0 * nr + t] = s_act[r * rw + net.r_off + j];
        out_v[n0 * nr + t] = s_act[r * rw + net.v_off + j];
      } else if (r == -2) {
        out_r[n0 * nr + t] = 0.f;
        out_v[n0 * nr + t] = 0.f;
      }
    }
#pragma unroll 4
    for (int t = tid; t < cn * nc; t += MOTION_THREADS) {
      const int i = t / nc, j = t - i * nc;
      const int r = s_slot[i];
      if (r >= 0) out_ccw[n0 * nc + t] = s_act[r * rw + net.w_off + j];
      else if (r == -2) out_ccw[n0 * nc + t] = j == 0 ? 1.f : 0.f;
    }
  }
}

__global__ void __launch_bounds__(MOTION_THREADS)
motion_bwd_kernel(const __grid_constant__ lush_motion_net net, int n_img, const int64_t* __restrict__ idx,
                  int64_t N, const float* __restrict__ rows, const float* __restrict__ g_r,
                  const float* __restrict__ g_v, const float* __restrict__ g_ccw,
                  float* __restrict__ g_embed, int stage_w) {
  extern __shared__ float s_mem[];
  __shared__ int s_slot[MOTION_CHUNK];
  const int rw = net.row_width, C = net.embed_dim, M = net.num_motion;
  float* s_act = s_mem;                                 // [MOTION_ROWS][rw]
  float* s_g = s_mem + MOTION_ROWS * rw;                // [MOTION_ROWS][rw]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = MOTION_THREADS / 32;
  const int row0 = blockIdx.x * MOTION_ROWS;
  const int nrows = min(MOTION_ROWS, n_img - row0);
  float* s_w = s_mem + 2 * MOTION_ROWS * rw;            // staged weights (stage_w)
  if (stage_w) motion_stage_weights(net, s_w, tid);
  for (int t = tid; t < MOTION_ROWS * rw; t += MOTION_THREADS) {
    s_act[t] = t < nrows * rw ? rows[(int64_t)row0 * rw + t] : 0.f;
    s_g[t] = 0.f;
  }
  __syncthreads();
  // per-image sums of the per-ray gradients: image slots staged per chunk of rays, then one thread per (ray, channel)
  // with shared-memory atomics (the lanes of a warp hit different channels of at most a few rays)
  const int nr = 3 * M, nc = M + 1;
  for (int64_t n0 = 0; n0 < N; n0 += MOTION_CHUNK) {
    const int cn = (int)min((int64_t)MOTION_CHUNK, N - n0);
    __syncthreads();
    for (int i = tid; i < cn; i += MOTION_THREADS) {
      const int64_t r = idx[n0 + i] - row0;
      s_slot[i] = (r >= 0 && r < nrows) ? (int)r : -1;
    }
    __syncthreads();
#pragma unroll 4
    for (int t = tid; t < cn * nr; t += MOTION_THREADS) {
      const int i = t / nr, j = t - i * nr;
      const int r = s_slot[i];
      if (r >= 0) {
        if (g_r) atomicAdd(&s_g[r * rw + net.r_off + j], g_r[n0 * nr + t]);
        if (g_v) atomicAdd(&s_g[r * rw + net.v_off + j], g_v[n0 * nr + t]);
      }
    }
    if (g_ccw) {
#pragma unroll 4
      for (int t = tid; t < cn * nc; t += MOTION_THREADS) {
        const int i = t / nc, j = t - i * nc;
        const int r = s_slot[i];
        if (r >= 0) atomicAdd(&s_g[r * rw + net.w_off + j], g_ccw[n0 * nc + t]);
      }
    }
  }
  __syncthreads();
  // heads backward
  for (int t = tid; t < MOTION_ROWS * 6 * M; t += MOTION_THREADS) {
    int r = t / (6 * M), j = t - r * 6 * M;
    int off = j < 3 * M ? net.r_off + j : net.v_off + (j - 3 * M);
    s_g[r * rw + off] *= net.rv_window;
  }
  if (tid < MOTION_ROWS) {
    const float* a = s_act + tid * rw + net.w_off;
    float* g = s_g + tid * rw + net.w_off;
    float sum = 1e-10f, dot = 0.f;
    for (int j = 0; j <= M; ++j) {
      float w = sigmoidf_(a[j]);
      sum += w;
      dot += g[j] * w;
    }
    for (int j = 0; j <= M; ++j) {
      float w = sigmoidf_(a[j]);
      float gw = g[j] / sum - dot / (sum * sum);
      g[j] = gw * w * (1.f - w);
    }
  }
  __syncthreads();
  int woff = motion_weight_offset(net, net.n_layers);     // running offset into the staged weights (descending)
  for (int l = net.n_layers - 1; l >= 0; --l) {
    const LayerRegs L = motion_layer(net, l);
    const int in_dim = L.in0_dim + L.in1_dim;
    woff -= L.out_dim * in_dim;
    if (L.act) {
      for (int t = tid; t < MOTION_ROWS * L.out_dim; t += MOTION_THREADS) {
        int r = t / L.out_dim, o = t - r * L.out_dim;
        if (s_act[r * rw + L.out_off + o] <= 0.f) s_g[r * rw + L.out_off + o] = 0.f;
      }
      __syncthreads();
    }
    // weight / bias gradients: warp per output unit, lanes over its fan-in (coalesced reductions into global)
    for (int o = warp; o < L.out_dim; o += nwarps) {
      float g[MOTION_ROWS], gsum = 0.f;
#pragma unroll
      for (int r = 0; r < MOTION_ROWS; ++r) {
        g[r] = s_g[r * rw + L.out_off + o];
        gsum += g[r];
      }
      float* gwrow = L.gw + (int64_t)o * in_dim;
      for (int k = lane; k < in_dim; k += 32) {
        const float* a = s_act + motion_src(L, k);
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < MOTION_ROWS; ++r) s = fmaf(g[r], a[r * rw], s);
        atomicAdd(&gwrow[k], s);
      }
      if (lane == 0) atomicAdd(&L.gb[o], gsum);
    }
    // input gradients: thread per input column, output units split over the thread groups
    if (l > 0 || g_embed) {
      const float* wsrc = stage_w ? s_w + woff : L.w;
      const int groups = max(1, MOTION_THREADS / in_dim);
      for (int k = tid % in_dim, og = tid / in_dim; og < groups && k < in_dim; k += MOTION_THREADS) {
        float acc[MOTION_ROWS];
#pragma unroll
        for (int r = 0; r < MOTION_ROWS; ++r) acc[r] = 0.f;
#pragma unroll 4
        for (int o = og; o < L.out_dim; o += groups) {
          float w = wsrc[(int64_t)o * in_dim + k];
          const float* g = s_g + L.out_off + o;
#pragma unroll
          for (int r = 0; r < MOTION_ROWS; ++r) acc[r] = fmaf(w, g[r * rw], acc[r]);
        }
        int src = motion_src(L, k);
#pragma unroll
        for (int r = 0; r < MOTION_ROWS; ++r) atomicAdd(&s_g[r * rw + src], acc[r]);
      }
    }
    __syncthreads();
  }
  if (g_embed)
    for (int t = tid; t < nrows * C; t += MOTION_THREADS) {
      int r = t / C, k = t - r * C;
      g_embed[(int64_t)(row0 + r) * C + k] = s_g[r * rw + k];
    }
}

static size_t motion_weight_bytes(const lush_motion_net* net) {
  size_t n = 0;
  for (int l = 0; l < net->n_layers; ++l)
    n += (size_t)net->layers[l].out_dim * (net->layers[l].in0_dim + net->layers[l].in1_dim + 1);   // + bias
  return n * sizeof(float);
}

static int motion_check(const lush_motion_net* net, const char* who) {
  LUSH_REQUIRE(net && net->n_layers >= 1 && net->n_layers <= LUSH_MOTION_MAX_LAYERS, "%s: bad layer count", who);
  LUSH_REQUIRE(net->embed_dim >= 1 && net->num_motion >= 1 && net->row_width >= net->embed_dim,
               "%s: bad network dimensions", who);
  const int rw = net->row_width, M = net->num_motion;
  for (int l = 0; l < net->n_layers; ++l) {
    const lush_dense_layer& L = net->layers[l];
    LUSH_REQUIRE(L.w && L.b && L.in0_dim >= 1 && L.in1_dim >= 0 && L.out_dim >= 1, "%s: layer %d incomplete", who, l);
    LUSH_REQUIRE(L.in0_off >= 0 && L.in0_off + L.in0_dim <= rw && L.out_off >= net->embed_dim &&
                     L.out_off + L.out_dim <= rw && (L.in1_dim == 0 || (L.in1_off >= 0 && L.in1_off + L.in1_dim <= rw)),
                 "%s: layer %d reads or writes outside the activation row", who, l);
    LUSH_REQUIRE(L.in0_off + L.in0_dim <= L.out_off && (L.in1_dim == 0 || L.in1_off + L.in1_dim <= L.out_off),
                 "%s: layer %d must read columns that precede its output", who, l);
  }
  LUSH_REQUIRE(net->r_off >= 0 && net->r_off + 3 * M <= rw && net->v_off >= 0 && net->v_off + 3 * M <= rw &&
                   net->w_off >= 0 && net->w_off + M + 1 <= rw,
               "%s: head offsets outside the activation row", who);
  return 0;
}

}  // namespace lush

using namespace lush;

extern "C" int lush_rbk_motion_fwd(const lush_motion_net* net, const float* embed, int n_img,
                                   const int64_t* images_idx, int64_t N, float* rows, float* r, float* v,
                                   float* ccw, int* err_flag, void* stream) {
  int rc = motion_check(net, "rbk_motion_fwd");
  if (rc) return rc;
  LUSH_REQUIRE(embed && n_img >= 1 && N >= 0 && rows && (N == 0 || (images_idx && r && v && ccw)),
               "rbk_motion_fwd: bad arguments");
  size_t smem = (size_t)MOTION_ROWS * net->row_width * sizeof(float);
  const size_t wbytes = motion_weight_bytes(net);
  const int stage_w = (smem + wbytes <= 180 * 1024) ? 1 : 0;
  if (stage_w) smem += wbytes;
  LUSH_REQUIRE(smem <= 200 * 1024, "rbk_motion_fwd: activation row of %d floats does not fit shared memory",
               net->row_width);
  {   // per call: function attributes are per device, and a process may drive several (DataParallel-style callers)
    cudaError_t e = cudaFuncSetAttribute(motion_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    LUSH_REQUIRE(e == cudaSuccess, "rbk_motion_fwd: %s", cudaGetErrorString(e));
  }
  unsigned blocks = (unsigned)ceil_div(n_img, MOTION_ROWS);
  motion_fwd_kernel<<<blocks, MOTION_THREADS, smem, (cudaStream_t)stream>>>(*net, embed, n_img, images_idx, N, rows,
                                                                            r, v, ccw, stage_w, err_flag);
  LUSH_LAUNCH_CHECK("rbk_motion_fwd");
  return 0;
}

extern "C" int lush_rbk_motion_bwd(const lush_motion_net* net, int n_img, const int64_t* images_idx, int64_t N,
                                   const float* rows, const float* g_r, const float* g_v, const float* g_ccw,
                                   float* g_embed, void* stream) {
  int rc = motion_check(net, "rbk_motion_bwd");
  if (rc) return rc;
  LUSH_REQUIRE(n_img >= 1 && N >= 0 && rows && (N == 0 || images_idx), "rbk_motion_bwd: bad arguments");
  for (int l = 0; l < net->n_layers; ++l)
    LUSH_REQUIRE(net->layers[l].gw && net->layers[l].gb, "rbk_motion_bwd: layer %d has no gradient buffers", l);
  size_t smem = (size_t)2 * MOTION_ROWS * net->row_width * sizeof(float);
  const size_t wbytes = motion_weight_bytes(net);
  const int stage_w = (smem + wbytes <= 180 * 1024) ? 1 : 0;
  if (stage_w) smem += wbytes;
  LUSH_REQUIRE(smem <= 200 * 1024, "rbk_motion_bwd: activation row of %d floats does not fit shared memory",
               net->row_width);
  {   // per call: function attributes are per device, and a process may drive several (DataParallel-style callers)
    cudaError_t e = cudaFuncSetAttribute(motion_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    LUSH_REQUIRE(e == cudaSuccess, "rbk_motion_bwd: %s", cudaGetErrorString(e));
  }
  unsigned blocks = (unsigned)ceil_div(n_img, MOTION_ROWS);
  motion_bwd_kernel<<<blocks, MOTION_THREADS, smem, (cudaStream_t)stream>>>(*net, n_img, images_idx, N, rows, g_r,
                                                                            g_v, g_ccw, g_embed, stage_w);
  LUSH_LAUNCH_CHECK("rbk_motion_bwd");
  return 0;
}
