// raw2outputs alpha compositing, forward and closed-form backward, as warp-per-ray segmented scans.
// Replaces NeRFAll.raw2outputs (reference models/lushnerf.py:296-352) and its autograd graph.
//
// Layout: raw [R,S,4] is read as one float4 per sample; lane l of the ray's warp owns samples
// l, l+32, l+64, ... so every load instruction of a warp covers 512 contiguous bytes of raw and 128
// contiguous bytes of z / noise / weights.  All loads of a ray are issued before the scan starts
// (NCH chunks held in registers) so each warp has NCH*3 independent requests in flight.
// HBM-bound: algorithmic bytes 32*S+24 per ray forward, 40*S+40 backward.
#include "common.cuh"

namespace lush {

constexpr int kWarpsPerBlock = 8;

// Colour activation.  1 / (1 + exp(-x)) with the hardware exp2 / reciprocal units (ex2.approx, rcp.approx: <= 2 ulp each):
// |error| <= 3e-7 on a value in [0, 1], against <= 1e-7 for the exact expf + correctly rounded reciprocal it replaces
// -- one third of the instructions of these issue-bound kernels.  The transmittance path (exp(-sigma dist), products)
// stays exact: its errors would be amplified by the 1 - exp() cancellation.
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

struct RayChunk {
  float4 rw;     // raw sample
  float z, zn;   // z_i, z_{i+1}
  float nz;      // sigma noise
};

// forward quantities of one sample
struct SampleFwd {
  float sigma;   // density after relu (+ eval near-plane mask)
  float pre;     // raw_sigma + noise (relu gate)
  float mask;    // eval near-plane mask (1 in training)
  float dist;
  float alpha;
  float om;      // 1 - alpha (what the reference multiplies into the transmittance)
  float e;       // exp(-sigma*dist): d alpha / d(sigma*dist), more accurate than 1 - alpha when alpha ~ 1
};

__device__ __forceinline__ SampleFwd sample_fwd(const RayChunk& c, bool has_next, float dnorm,
                                                float thresh) {
  SampleFwd f;
  f.dist = (c.zn - c.z) * dnorm;
  f.pre = c.rw.w + c.nz;
  f.mask = 1.f;
  if (thresh >= 0.f) f.mask = (c.zn > thresh) ? 1.f : 0.f;
  f.sigma = fmaxf(f.pre, 0.f) * f.mask;
  if (has_next) {
    f.e = expf(-(f.sigma * f.dist));
    f.alpha = 1.f - f.e;
  } else {  // last sample: alpha = 1 (reference :338), no density entry
    f.alpha = 1.f;
    f.sigma = 0.f;
    f.dist = 0.f;
    f.e = 0.f;
  }
  f.om = 1.f - f.alpha;  // reference adds 1e-10 in fp32, which rounds away (-alpha + 1.0f)
  return f;
}

// row pointers of one ray (computed once: the per-sample address arithmetic was ~10 % of the instructions)
struct RayPtrs {
  const float4* raw;
  const float* z;
  const float* nz;   // nullptr: no sigma noise
};
__device__ __forceinline__ RayPtrs ray_ptrs(const float* raw, const float* z, const float* noise, int64_t r, int S) {
  RayPtrs p;
  p.raw = reinterpret_cast<const float4*>(raw) + r * S;
  p.z = z + r * S;
  p.nz = noise ? noise + r * (S - 1) : nullptr;
  return p;
}
// sample i of the ray: everything except z_{i+1} (see link_next)
__device__ __forceinline__ void load_chunk(RayChunk& c, const RayPtrs& p, int S, int i) {
  if (i < S) {
    c.rw = ld_stream4(p.raw + i);
    c.z = p.z[i];
    c.nz = (p.nz != nullptr && i + 1 < S) ? ld_stream(p.nz + i) : 0.f;
  } else {
    c.rw = make_float4(0.f, 0.f, 0.f, 0.f);
    c.z = c.nz = 0.f;
  }
  c.zn = c.z;
}
// z_{i+1} of the samples of chunk k: the next lane's z, lane 31 takes lane 0's z of chunk k + 1 (z_first_next)
__device__ __forceinline__ void link_next(RayChunk& c, float z_first_next, bool has_next, int lane) {
  float zn = __shfl_down_sync(0xffffffffu, c.z, 1);
  if (lane == 31) zn = z_first_next;
  c.zn = has_next ? zn : c.z;
}

// exclusive product scan over the 32 lanes; returns exclusive prefix, `total` = product of all lanes
__device__ __forceinline__ float warp_excl_prod(float v, int lane, float& total) {
  float incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl *= t;
  }
  total = __shfl_sync(0xffffffffu, incl, 31);
  float ex = __shfl_up_sync(0xffffffffu, incl, 1);
  return lane == 0 ? 1.f : ex;
}

// 1 iff the exponent field of v is all ones (NaN or Inf): two integer instructions; the exact NaN/Inf split is
// only computed for values that trip this test
__device__ __forceinline__ unsigned nonfinite_bit(float v) {
  return ((__float_as_uint(v) & 0x7fffffffu) + 0x00800000u) >> 31;
}
__device__ __forceinline__ int nanflag(float v) {
  return (isnan(v) ? 1 : 0) | (isinf(v) ? 2 : 0);
}

// NCH = number of 32-sample chunks (compile time), 0 = runtime loop (S > 256)
template <int NCH>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
composite_fwd_kernel(const float* __restrict__ raw, const float* __restrict__ z,
                     const float* __restrict__ rays_d, int d_stride,
                     const float* __restrict__ noise, float thresh, int white,
                     float* __restrict__ rgb_map, float* __restrict__ density,
                     float* __restrict__ acc_map, float* __restrict__ weights,
                     float* __restrict__ depth_map, int* __restrict__ flags, int64_t R, int S) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= R) return;
  const float dx = rays_d[r * d_stride], dy = rays_d[r * d_stride + 1], dz = rays_d[r * d_stride + 2];
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
  constexpr int N = NCH > 0 ? NCH : 1;
  const int nch = NCH > 0 ? NCH : (S + 31) / 32;
  const RayPtrs rp = ray_ptrs(raw, z, noise, r, S);
  float* __restrict__ w_row = weights ? weights + r * S : nullptr;
  float* __restrict__ d_row = density ? density + r * (S - 1) : nullptr;
  RayChunk ch[N];
  if (NCH > 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) load_chunk(ch[k], rp, S, k * 32 + lane);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const float zf = __shfl_sync(0xffffffffu, ch[k + 1 < N ? k + 1 : k].z, 0);
      link_next(ch[k], zf, k * 32 + lane + 1 < S, lane);
    }
  }
  float T_carry = 1.f, sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f, sa = 0.f;
  int fl = 0;
#pragma unroll
  for (int k = 0; k < nch; ++k) {
    const int i = k * 32 + lane;
    RayChunk c;
    if (NCH > 0) c = ch[k];
    else {      // runtime chunk count (S > 256): z_{i+1} straight from memory
      load_chunk(c, rp, S, i);
      if (i + 1 < S) c.zn = rp.z[i + 1];
    }
    const bool valid = i < S;
    SampleFwd f = sample_fwd(c, i + 1 < S, dnorm, thresh);
    float om = valid ? f.om : 1.f, tot;
    float T = T_carry * warp_excl_prod(om, lane, tot);
    T_carry *= tot;
    float w = valid ? f.alpha * T : 0.f;
    sr += w * sigmoid_fast(c.rw.x);
    sg += w * sigmoid_fast(c.rw.y);
    sb += w * sigmoid_fast(c.rw.z);
    sd += w * c.z;
    sa += w;
    if (valid) {
      if (w_row) w_row[i] = w;
      if (d_row && i + 1 < S) d_row[i] = f.sigma;
      if (nonfinite_bit(c.rw.x) | nonfinite_bit(c.rw.y) | nonfinite_bit(c.rw.z) | nonfinite_bit(c.rw.w) |
          nonfinite_bit(f.sigma)) {   // rare
        fl |= (nanflag(c.rw.x) | nanflag(c.rw.y) | nanflag(c.rw.z) | nanflag(c.rw.w)) << 8;
        if (i + 1 < S) fl |= nanflag(f.sigma) << 6;
      }
    }
  }
  sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb); sd = warp_sum(sd); sa = warp_sum(sa);
  if (white) { sr += 1.f - sa; sg += 1.f - sa; sb += 1.f - sa; }
  if (lane == 0) {
    rgb_map[r * 3] = sr; rgb_map[r * 3 + 1] = sg; rgb_map[r * 3 + 2] = sb;
    depth_map[r] = sd;
    acc_map[r] = sa;
    fl |= nanflag(sr) | nanflag(sg) | nanflag(sb) | (nanflag(sd) << 2) | (nanflag(sa) << 4);
  }
  if (flags != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fl |= __shfl_xor_sync(0xffffffffu, fl, o);
    if (fl != 0 && lane == 0) atomicOr(flags, fl);
  }
}

// ------------------------------------------------------------------------------------------------
// backward:  dL/dalpha_i = T_i (G_i - B_i),  B_i = G_{i+1} a_{i+1} + (1-a_{i+1}) B_{i+1},  B_{S-1}=0
// (division-free form of the closed form in SURVEY.md 8 a-6; exact also when a sample saturates)
// ------------------------------------------------------------------------------------------------
// occupancy floor per instantiation: the kernel hides its shuffle / exp latency chains only with >= 16 warps per SM
template <int NCH>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, (NCH <= 2 ? 1 : (NCH <= 4 ? 4 : (NCH <= 8 ? 2 : 1))))
composite_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ z,
                     const float* __restrict__ rays_d, int d_stride,
                     const float* __restrict__ noise, float thresh, int white,
                     const float* __restrict__ g_rgb, const float* __restrict__ g_density,
                     const float* __restrict__ g_acc, const float* __restrict__ g_weights,
                     const float* __restrict__ g_depth, float* __restrict__ g_raw,
                     float* __restrict__ g_rays_d, int64_t R, int S) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= R) return;
  const float dx = rays_d[r * d_stride], dy = rays_d[r * d_stride + 1], dz = rays_d[r * d_stride + 2];
  const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
  constexpr int N = NCH;
  const RayPtrs rp = ray_ptrs(raw, z, noise, r, S);
  float4* __restrict__ g_row = reinterpret_cast<float4*>(g_raw) + r * S;
  RayChunk ch[N];
  float gw_in[N], gden_in[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const int i = k * 32 + lane;
    load_chunk(ch[k], rp, S, i);
    gw_in[k] = (g_weights != nullptr && i < S) ? ld_stream(g_weights + r * S + i) : 0.f;
    gden_in[k] = (g_density != nullptr && i + 1 < S) ? ld_stream(g_density + r * (S - 1) + i) : 0.f;
  }
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const float zf = __shfl_sync(0xffffffffu, ch[k + 1 < N ? k + 1 : k].z, 0);
    link_next(ch[k], zf, k * 32 + lane + 1 < S, lane);
  }
  float gr = 0.f, gg = 0.f, gb = 0.f, gd = 0.f, ga = 0.f;
  if (g_rgb) { gr = g_rgb[r * 3]; gg = g_rgb[r * 3 + 1]; gb = g_rgb[r * 3 + 2]; }
  if (g_depth) gd = g_depth[r];
  if (g_acc) ga = g_acc[r];
  if (white) ga -= gr + gg + gb;  // rgb_map += 1 - acc

  // forward recompute: T_i for every owned sample
  float Tk[N];
  float T_carry = 1.f;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const int i = k * 32 + lane;
    SampleFwd f = sample_fwd(ch[k], i + 1 < S, dnorm, thresh);
    float om = (i < S) ? f.om : 1.f, tot;
    Tk[k] = T_carry * warp_excl_prod(om, lane, tot);
    T_carry *= tot;
  }
  // reverse pass
  float B_carry = 0.f;   // B at the last sample of the current chunk
  float gdn = 0.f;       // d L / d |rays_d|
#pragma unroll
  for (int k = N - 1; k >= 0; --k) {
    const int i = k * 32 + lane;
    const bool valid = i < S;
    const RayChunk& c = ch[k];
    SampleFwd f = sample_fwd(c, i + 1 < S, dnorm, thresh);
    float cr = sigmoid_fast(c.rw.x), cg = sigmoid_fast(c.rw.y), cb = sigmoid_fast(c.rw.z);
    float G = gr * cr + gg * cg + gb * cb + gd * c.z + ga + gw_in[k];
    // affine map f_i(x) = b + a x, identity on padding lanes
    float a = valid ? f.om : 1.f;
    float b = valid ? G * f.alpha : 0.f;
    // inclusive suffix composition F_i = f_i o f_{i+1} o ... o f_{chunk end}
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float a2 = __shfl_down_sync(0xffffffffu, a, o);
      float b2 = __shfl_down_sync(0xffffffffu, b, o);
      if (lane + o < 32) { b = b + a * b2; a = a * a2; }
    }
    // B_i = F_{i+1}(B_carry); lane 31: B_carry itself
    float an = __shfl_down_sync(0xffffffffu, a, 1);
    float bn = __shfl_down_sync(0xffffffffu, b, 1);
    float Bi = (lane == 31) ? B_carry : bn + an * B_carry;
    float a0 = __shfl_sync(0xffffffffu, a, 0), b0 = __shfl_sync(0xffffffffu, b, 0);
    B_carry = b0 + a0 * B_carry;  // B at sample (chunk start - 1)
    if (valid) {
      float T = Tk[k];
      float w = f.alpha * T;
      float4 go;
      go.x = w * gr * cr * (1.f - cr);
      go.y = w * gg * cg * (1.f - cg);
      go.z = w * gb * cb * (1.f - cb);
      go.w = 0.f;
      if (i + 1 < S) {
        float galpha = T * (G - Bi);
        float gsig = galpha * f.dist * f.e + gden_in[k];
        go.w = (f.pre > 0.f) ? gsig * f.mask : 0.f;
        gdn += galpha * f.sigma * f.e * (c.zn - c.z);
      }
      g_row[i] = go;
    }
  }
  if (g_rays_d != nullptr) {
    gdn = warp_sum(gdn);
    if (lane == 0) {
      float s = dnorm > 0.f ? gdn / dnorm : 0.f;
      g_rays_d[r * 3] = s * dx; g_rays_d[r * 3 + 1] = s * dy; g_rays_d[r * 3 + 2] = s * dz;
    }
  }
}

}  // namespace lush

using namespace lush;

#define DISPATCH_NCH(nch, CALL)            \
  switch (nch) {                           \
    case 1: { constexpr int NC = 1; CALL; } break; \
    case 2: { constexpr int NC = 2; CALL; } break; \
    case 3: { constexpr int NC = 3; CALL; } break; \
    case 4: { constexpr int NC = 4; CALL; } break; \
    case 5: { constexpr int NC = 5; CALL; } break; \
    case 6: { constexpr int NC = 6; CALL; } break; \
    case 7: { constexpr int NC = 7; CALL; } break; \
    case 8: { constexpr int NC = 8; CALL; } break; \
    default: { constexpr int NC = 0; CALL; } break; \
  }

extern "C" int lush_composite_fwd(const float* raw, const float* z, const float* rays_d,
                                  int d_stride, const float* sigma_noise, float rm_near_thresh,
                                  int white_bkgd, float* rgb_map, float* density, float* acc,
                                  float* weights, float* depth, int* flags, int64_t R, int S,
                                  void* stream) {
  LUSH_REQUIRE(raw && z && rays_d && rgb_map && acc && depth, "composite_fwd: null pointer");
  LUSH_REQUIRE(R >= 0 && S >= 1 && d_stride >= 3, "composite_fwd: bad R/S/d_stride");
  LUSH_REQUIRE(((uintptr_t)raw & 15) == 0, "composite_fwd: raw must be 16-byte aligned");
  if (R == 0) return 0;
  dim3 grid((unsigned)ceil_div(R, kWarpsPerBlock)), block(kWarpsPerBlock * 32);
  cudaStream_t st = (cudaStream_t)stream;
  int nch = (S + 31) / 32;
  DISPATCH_NCH(nch, (composite_fwd_kernel<NC><<<grid, block, 0, st>>>(
                        raw, z, rays_d, d_stride, sigma_noise, rm_near_thresh, white_bkgd, rgb_map,
                        density, acc, weights, depth, flags, R, S)));
  LUSH_LAUNCH_CHECK("composite_fwd");
  return 0;
}

extern "C" int lush_composite_bwd(const float* raw, const float* z, const float* rays_d,
                                  int d_stride, const float* sigma_noise, float rm_near_thresh,
                                  int white_bkgd, const float* g_rgb, const float* g_density,
                                  const float* g_acc, const float* g_weights, const float* g_depth,
                                  float* g_raw, float* g_rays_d, int64_t R, int S, void* stream) {
  LUSH_REQUIRE(raw && z && rays_d && g_raw, "composite_bwd: null pointer");
  LUSH_REQUIRE(R >= 0 && S >= 1 && d_stride >= 3, "composite_bwd: bad R/S/d_stride");
  LUSH_REQUIRE(S <= 512, "composite_bwd: S > 512 samples per ray is not supported");
  LUSH_REQUIRE((((uintptr_t)raw | (uintptr_t)g_raw) & 15) == 0,
               "composite_bwd: raw/g_raw must be 16-byte aligned");
  if (R == 0) return 0;
  dim3 grid((unsigned)ceil_div(R, kWarpsPerBlock)), block(kWarpsPerBlock * 32);
  cudaStream_t st = (cudaStream_t)stream;
  int nch = (S + 31) / 32;
#define BWD_CALL(NCV)                                                                            \
  composite_bwd_kernel<NCV><<<grid, block, 0, st>>>(raw, z, rays_d, d_stride, sigma_noise,       \
                                                    rm_near_thresh, white_bkgd, g_rgb, g_density, \
                                                    g_acc, g_weights, g_depth, g_raw, g_rays_d, R, S)
  if (nch <= 1) BWD_CALL(1);
  else if (nch <= 2) BWD_CALL(2);
  else if (nch <= 3) BWD_CALL(3);
  else if (nch <= 4) BWD_CALL(4);
  else if (nch <= 6) BWD_CALL(6);
  else if (nch <= 8) BWD_CALL(8);
  else if (nch <= 12) BWD_CALL(12);
  else BWD_CALL(16);
#undef BWD_CALL
  LUSH_LAUNCH_CHECK("composite_bwd");
  return 0;
}
