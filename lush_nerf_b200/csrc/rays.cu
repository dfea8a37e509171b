// Per-ray elementwise kernels: ray preparation (normalise + NDC), coarse sample placement,
// rigid-warp blur kernel (SE(3) exp map) forward/backward, K-ray blend + tone map forward/backward.
// All are tiny, HBM/latency-bound; one thread per ray (or pixel).  Compiled with --fmad=false so the
// fp32 results follow the reference's unfused op sequence.
#include "common.cuh"

namespace lush {

struct NdcConst {
  float cx, cy;  // -1/(W/(2f)), -1/(H/(2f))  (reference helpers:551-556, Python double -> fp32)
};

// ------------------------------------------------------------------------------------------------
// ray_prep  (reference models/lushnerf.py:706-729 + helpers:542-562)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ray_prep_store(float ox, float oy, float oz, float dx, float dy, float dz, NdcConst k,
                                               float near, float far, int ndc, float* __restrict__ q) {
  float n = sqrtf(dx * dx + dy * dy + dz * dz);
  float vx = dx / n, vy = dy / n, vz = dz / n;
  float o0 = ox, o1 = oy, o2 = oz, d0 = dx, d1 = dy, d2 = dz;
  if (ndc) {
    float t = -(1.f + oz) / dz;
    float px = ox + t * dx, py = oy + t * dy, pz = oz + t * dz;
    o0 = k.cx * px / pz;
    o1 = k.cy * py / pz;
    o2 = 1.f + 2.f / pz;
    d0 = k.cx * (dx / dz - px / pz);
    d1 = k.cy * (dy / dz - py / pz);
    d2 = -2.f / pz;
  }
  q[0] = o0; q[1] = o1; q[2] = o2; q[3] = d0; q[4] = d1; q[5] = d2;
  q[6] = near; q[7] = far; q[8] = vx; q[9] = vy; q[10] = vz;
}

__global__ void ray_prep_fwd_kernel(const float* __restrict__ rays, float* __restrict__ out,
                                    int64_t R, NdcConst k, float near, float far, int ndc) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float* p = rays + r * 6;
  ray_prep_store(p[0], p[2], p[4], p[1], p[3], p[5], k, near, far, ndc, out + r * 11);
}

// Rays of pixels [pix0, pix0 + n) of one pose, generated on the fly and prepared in the same pass: get_rays
// (reference helpers:517-528: dirs = ((i + .5 - cx)/fx, -(j + .5 - cy)/fy, -1), d = sum_k dirs_k c2w[:, k],
// o = c2w[:, 3]) fused with the ray preparation above -- the [H, W, 3, 2] ray image of render_path
// (models/lushnerf.py:881-882) is never materialised.
struct PinholeConst {
  float ax, ay, fx, fy;   // (.5 - cx), (.5 - cy) formed in double like the Python expression, then fp32
};
__global__ void ray_gen_prep_kernel(const float* __restrict__ c2w, int c2w_stride, PinholeConst cam, int W,
                                    int64_t pix0, int64_t n, float* __restrict__ out, NdcConst k, float near,
                                    float far, int ndc) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int64_t p = pix0 + t;
  const int64_t j = p / W;
  const float fi = (float)(p - j * W), fj = (float)j;
  const float u = (fi + cam.ax) / cam.fx, v = -((fj + cam.ay) / cam.fy), w = -1.f;
  float d[3], o[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float* row = c2w + r * c2w_stride;
    d[r] = (u * row[0] + v * row[1]) + w * row[2];
    o[r] = row[3];
  }
  ray_prep_store(o[0], o[1], o[2], d[0], d[1], d[2], k, near, far, ndc, out + t * 11);
}

__global__ void ray_prep_bwd_kernel(const float* __restrict__ rays, const float* __restrict__ g,
                                    float* __restrict__ g_rays, int64_t R, NdcConst k, int ndc) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float* p = rays + r * 6;
  float ox = p[0], dx = p[1], oy = p[2], dy = p[3], oz = p[4], dz = p[5];
  const float* gq = g + r * 11;
  float go0 = gq[0], go1 = gq[1], go2 = gq[2], gd0 = gq[3], gd1 = gq[4], gd2 = gq[5];
  float gv0 = gq[8], gv1 = gq[9], gv2 = gq[10];
  float gox, goy, goz, gdx, gdy, gdz;
  if (ndc) {
    float t = -(1.f + oz) / dz;
    float px = ox + t * dx, py = oy + t * dy, pz = oz + t * dz;
    float a = px / pz, b = py / pz;
    float ga = k.cx * (go0 - gd0), gb = k.cy * (go1 - gd1);
    float gpx = ga / pz, gpy = gb / pz;
    float gpz = (-ga * a - gb * b) / pz + 2.f * (gd2 - go2) / (pz * pz);
    gdx = k.cx * gd0 / dz;
    gdy = k.cy * gd1 / dz;
    gdz = -(k.cx * gd0 * dx + k.cy * gd1 * dy) / (dz * dz);
    // p = o + t d
    float gt = gpx * dx + gpy * dy + gpz * dz;
    gox = gpx; goy = gpy; goz = gpz;
    gdx += t * gpx; gdy += t * gpy; gdz += t * gpz;
    // t = -(1+oz)/dz
    goz += -gt / dz;
    gdz += gt * (1.f + oz) / (dz * dz);
  } else {
    gox = go0; goy = go1; goz = go2; gdx = gd0; gdy = gd1; gdz = gd2;
  }
  // viewdirs = d/|d|
  float n = sqrtf(dx * dx + dy * dy + dz * dz);
  float vx = dx / n, vy = dy / n, vz = dz / n;
  float dot = vx * gv0 + vy * gv1 + vz * gv2;
  gdx += (gv0 - vx * dot) / n;
  gdy += (gv1 - vy * dot) / n;
  gdz += (gv2 - vz * dot) / n;
  float* q = g_rays + r * 6;
  q[0] = gox; q[1] = gdx; q[2] = goy; q[3] = gdy; q[4] = goz; q[5] = gdz;
}

// ------------------------------------------------------------------------------------------------
// stratified z  (reference models/lushnerf.py:389-412)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float z_at(float near, float far, float t, int lindisp) {
  if (!lindisp) return near * (1.f - t) + far * t;
  return 1.f / (1.f / near * (1.f - t) + 1.f / far * t);
}

__global__ void stratified_z_kernel(const float* __restrict__ ray_batch, int ray_stride,
                                    const float* __restrict__ t_vals,
                                    const float* __restrict__ t_rand, float* __restrict__ z_out,
                                    int64_t R, int S, int lindisp) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= R * S) return;
  int64_t r = idx / S;
  int s = (int)(idx - r * S);
  float near = ray_batch[r * ray_stride + 6], far = ray_batch[r * ray_stride + 7];
  float zc = z_at(near, far, t_vals[s], lindisp);
  if (t_rand != nullptr) {
    float lower = zc, upper = zc;
    if (s > 0) lower = 0.5f * (zc + z_at(near, far, t_vals[s - 1], lindisp));
    if (s < S - 1) upper = 0.5f * (z_at(near, far, t_vals[s + 1], lindisp) + zc);
    zc = lower + (upper - lower) * t_rand[idx];
  }
  z_out[idx] = zc;
}

// ------------------------------------------------------------------------------------------------
// rigid-warp blur kernel (reference models/lushnerf.py:75-98, utils/rigid_warping.py:20-140)
// ------------------------------------------------------------------------------------------------
struct V3 {
  float x, y, z;
};
__device__ __forceinline__ V3 mk(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

struct Se3 {
  V3 w, u;           // unit axis, scaled translation
  float th, s, c1, k;  // theta, sin, 1-cos, theta-sin
  V3 au, bu, p;
};

__device__ __forceinline__ Se3 se3_make(V3 rot, V3 trans) {
  Se3 e;
  float nr = sqrtf(dot(rot, rot));
  e.th = nr + 1.0e-10f;
  e.w = mk(rot.x / e.th, rot.y / e.th, rot.z / e.th);
  e.u = mk(trans.x / e.th, trans.y / e.th, trans.z / e.th);
  float sn, cs;
  sincosf(e.th, &sn, &cs);
  e.s = sn;
  e.c1 = 1.f - cs;
  e.k = e.th - sn;
  e.au = cross(e.w, e.u);
  e.bu = cross(e.w, e.au);
  e.p = e.th * e.u + e.c1 * e.au + e.k * e.bu;
  return e;
}
__device__ __forceinline__ V3 se3_apply(const Se3& e, V3 x) {
  V3 a = cross(e.w, x);
  V3 b = cross(e.w, a);
  return x + e.s * a + e.c1 * b + e.p;
}

__global__ void rbk_warp_fwd_kernel(const float* __restrict__ rays, const float* __restrict__ r,
                                    const float* __restrict__ v, float* __restrict__ out, int64_t N,
                                    int M) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= N * (M + 1)) return;
  int64_t n = idx / (M + 1);
  int m = (int)(idx - n * (M + 1));
  const float* p = rays + n * 6;
  V3 o = mk(p[0], p[2], p[4]), d = mk(p[1], p[3], p[5]);
  float* q = out + idx * 6;
  if (m == 0) {
    q[0] = o.x; q[1] = d.x; q[2] = o.y; q[3] = d.y; q[4] = o.z; q[5] = d.z;
    return;
  }
  int mm = m - 1;
  const float* rp = r + n * 3 * M;
  const float* vp = v + n * 3 * M;
  Se3 e = se3_make(mk(rp[mm], rp[M + mm], rp[2 * M + mm]), mk(vp[mm], vp[M + mm], vp[2 * M + mm]));
  V3 yo = se3_apply(e, o);
  V3 ye = se3_apply(e, o + d);
  V3 yd = ye - yo;
  q[0] = yo.x; q[1] = yd.x; q[2] = yo.y; q[3] = yd.y; q[4] = yo.z; q[5] = yd.z;
}

// reverse of y = x + s (w x x) + c1 (w x (w x x)) for upstream g; accumulates gw, gs, gc1, returns gx
__device__ __forceinline__ V3 rot_apply_bwd(const Se3& e, V3 x, V3 g, V3& gw, float& gs, float& gc1) {
  V3 a = cross(e.w, x);
  V3 b = cross(e.w, a);
  gs += dot(g, a);
  gc1 += dot(g, b);
  V3 ga = e.s * g;
  V3 gb = e.c1 * g;
  // b = w x a
  gw = gw + cross(a, gb);
  ga = ga + cross(gb, e.w);
  // a = w x x
  gw = gw + cross(x, ga);
  return g + cross(ga, e.w);
}

__global__ void rbk_warp_bwd_kernel(const float* __restrict__ rays, const float* __restrict__ r,
                                    const float* __restrict__ v, const float* __restrict__ g_out,
                                    float* __restrict__ g_rays, float* __restrict__ g_r,
                                    float* __restrict__ g_v, int64_t N, int M) {
  int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* p = rays + n * 6;
  V3 o = mk(p[0], p[2], p[4]), d = mk(p[1], p[3], p[5]);
  const float* g0 = g_out + n * (M + 1) * 6;
  V3 go = mk(g0[0], g0[2], g0[4]), gd = mk(g0[1], g0[3], g0[5]);
  const float* rp = r + n * 3 * M;
  const float* vp = v + n * 3 * M;
  for (int mm = 0; mm < M; ++mm) {
    const float* gq = g0 + (mm + 1) * 6;
    V3 gyo = mk(gq[0], gq[2], gq[4]), gyd = mk(gq[1], gq[3], gq[5]);
    V3 rot = mk(rp[mm], rp[M + mm], rp[2 * M + mm]);
    V3 trans = mk(vp[mm], vp[M + mm], vp[2 * M + mm]);
    Se3 e = se3_make(rot, trans);
    V3 gye = gyd;
    V3 gyo_t = gyo - gyd;
    V3 gp = gyo;  // = gye + gyo_t
    V3 gw = mk(0, 0, 0);
    float gs = 0.f, gc1 = 0.f;
    V3 gx_o = rot_apply_bwd(e, o, gyo_t, gw, gs, gc1);
    V3 gx_e = rot_apply_bwd(e, o + d, gye, gw, gs, gc1);
    go = go + gx_o + gx_e;
    gd = gd + gx_e;
    // p = th*u + c1*au + k*bu
    float gth = dot(gp, e.u);
    V3 gu = e.th * gp;
    gc1 += dot(gp, e.au);
    float gk = dot(gp, e.bu);
    V3 gau = e.c1 * gp;
    V3 gbu = e.k * gp;
    gw = gw + cross(e.au, gbu);
    gau = gau + cross(gbu, e.w);
    gw = gw + cross(e.u, gau);
    gu = gu + cross(gau, e.w);
    // k = th - s ; s = sin th ; c1 = 1 - cos th
    gth += gk;
    gs -= gk;
    float cs = 1.f - e.c1;
    gth += gs * cs + gc1 * e.s;
    // w = rot/th ; u = trans/th
    V3 grot = (1.f / e.th) * gw;
    V3 gtr = (1.f / e.th) * gu;
    gth += -(dot(gw, e.w) + dot(gu, e.u)) / e.th;
    // th = |rot| + eps
    float nr = sqrtf(dot(rot, rot));
    if (nr > 0.f) grot = grot + (gth / nr) * rot;
    float* grp = g_r + n * 3 * M;
    float* gvp = g_v + n * 3 * M;
    grp[mm] = grot.x; grp[M + mm] = grot.y; grp[2 * M + mm] = grot.z;
    gvp[mm] = gtr.x;  gvp[M + mm] = gtr.y;  gvp[2 * M + mm] = gtr.z;
  }
  float* q = g_rays + n * 6;
  q[0] = go.x; q[1] = gd.x; q[2] = go.y; q[3] = gd.y; q[4] = go.z; q[5] = gd.z;
}

// ------------------------------------------------------------------------------------------------
// K-ray blend + noise + tone map (reference models/lushnerf.py:100-116, 649-654; helpers:164-174)
// ------------------------------------------------------------------------------------------------
constexpr float kInvGamma = (float)(1.0 / 2.2);

__global__ void blend_tonemap_fwd_kernel(const float* __restrict__ rgb, const float* __restrict__ ccw,
                                         const float* __restrict__ noise_raw,
                                         float* __restrict__ out_blur, float* __restrict__ out_pure,
                                         float* __restrict__ noise_rgb, int64_t N, int K, int gamma) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= N * 3) return;
  int64_t n = idx / 3;
  int c = (int)(idx - n * 3);
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    float wk = ccw ? ccw[n * K + k] : 1.f;
    acc += rgb[(n * K + k) * 3 + c] * wk;
  }
  float nz = 0.f;
  if (noise_raw) {
    nz = 0.1f * sigmoidf_(noise_raw[idx]);
    if (noise_rgb) noise_rgb[idx] = nz;
  }
  float xb = acc + nz;
  if (out_blur) out_blur[idx] = gamma ? powf(xb, kInvGamma) : xb;
  if (out_pure) out_pure[idx] = gamma ? powf(acc, kInvGamma) : acc;
}

__global__ void blend_tonemap_bwd_kernel(const float* __restrict__ rgb, const float* __restrict__ ccw,
                                         const float* __restrict__ noise_raw,
                                         const float* __restrict__ g_blur,
                                         const float* __restrict__ g_pure,
                                         const float* __restrict__ g_noise_rgb,
                                         float* __restrict__ g_rgb, float* __restrict__ g_ccw,
                                         float* __restrict__ g_noise_raw, int64_t N, int K, int gamma) {
  int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (n >= N) return;
  float gacc[3];
  for (int c = 0; c < 3; ++c) {
    int64_t idx = n * 3 + c;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc += rgb[(n * K + k) * 3 + c] * (ccw ? ccw[n * K + k] : 1.f);
    float sg = 0.f, nz = 0.f;
    if (noise_raw) {
      sg = sigmoidf_(noise_raw[idx]);
      nz = 0.1f * sg;
    }
    float gb = g_blur ? g_blur[idx] : 0.f;
    float gp = g_pure ? g_pure[idx] : 0.f;
    float gxb = gb, gxp = gp;
    if (gamma) {
      gxb = gb * kInvGamma * powf(acc + nz, kInvGamma - 1.f);
      gxp = gp * kInvGamma * powf(acc, kInvGamma - 1.f);
      if (gb == 0.f) gxb = 0.f;
      if (gp == 0.f) gxp = 0.f;
    }
    gacc[c] = gxb + gxp;
    if (g_noise_raw) {
      float gn = gxb + (g_noise_rgb ? g_noise_rgb[idx] : 0.f);
      g_noise_raw[idx] = noise_raw ? gn * 0.1f * sg * (1.f - sg) : 0.f;
    }
  }
  for (int k = 0; k < K; ++k) {
    float wk = ccw ? ccw[n * K + k] : 1.f;
    float gw = 0.f;
    for (int c = 0; c < 3; ++c) {
      int64_t j = (n * K + k) * 3 + c;
      if (g_rgb) g_rgb[j] = gacc[c] * wk;
      gw += gacc[c] * rgb[j];
    }
    if (g_ccw) g_ccw[n * K + k] = gw;
  }
}

}  // namespace lush

using namespace lush;

static NdcConst make_ndc(int H, int W, double focal) {
  NdcConst k;
  k.cx = (float)(-1.0 / (W / (2.0 * focal)));
  k.cy = (float)(-1.0 / (H / (2.0 * focal)));
  return k;
}

extern "C" int lush_ray_prep_fwd(const float* rays, float* ray_batch, int64_t R, int H, int W,
                                 double focal, float near, float far, int ndc, void* stream) {
  LUSH_REQUIRE(rays && ray_batch && R >= 0, "ray_prep_fwd: null pointer or negative R");
  if (R == 0) return 0;
  int th = 256;
  ray_prep_fwd_kernel<<<(unsigned)ceil_div(R, th), th, 0, (cudaStream_t)stream>>>(
      rays, ray_batch, R, make_ndc(H, W, focal), near, far, ndc);
  LUSH_LAUNCH_CHECK("ray_prep_fwd");
  return 0;
}

namespace lush {
// Raw rays (o, d) of arbitrary (view, pixel) pairs, e.g. a training batch drawn on the device: replaces the
// pre-expanded [V, H, W, ro+rd+rgb, 3] ray table and its host-side fancy indexing (reference run_lushnerf.py:538-616,
// get_rays_np helpers:531-539).  rays [n, 3, 2] in the layout NeRFAll.forward takes.
__global__ void ray_gen_gather_kernel(const float* __restrict__ poses, int n_views, int pose_stride, int row_stride,
                                      const int64_t* __restrict__ view, const int64_t* __restrict__ pix,
                                      PinholeConst cam, int W, int64_t npix, int64_t n, float* __restrict__ rays,
                                      const float* __restrict__ images, float* __restrict__ rgb_out,
                                      const uint8_t* __restrict__ masks, uint8_t* __restrict__ mask_out,
                                      int* __restrict__ err_flag) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n) return;
  int64_t vi = view[t], p = pix[t];
  int bad = 0;
  if (vi < 0 || vi >= n_views) { bad |= LUSH_FLAG_BAD_VIEW_INDEX; vi = vi < 0 ? 0 : n_views - 1; }
  if (p < 0 || p >= npix) { bad |= LUSH_FLAG_BAD_PIXEL_INDEX; p = p < 0 ? 0 : npix - 1; }
  if (bad && err_flag) atomicOr(err_flag, bad);     // rare
  const float* c2w = poses + vi * pose_stride;
  const int64_t j = p / W;
  const float fi = (float)(p - j * W), fj = (float)j;
  const float u = (fi + cam.ax) / cam.fx, v = -((fj + cam.ay) / cam.fy), w = -1.f;
  float* q = rays + t * 6;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float* row = c2w + r * row_stride;
    q[2 * r] = row[3];
    q[2 * r + 1] = (u * row[0] + v * row[1]) + w * row[2];
  }
  if (images) {
    const float* src = images + (vi * npix + p) * 3;
    rgb_out[t * 3 + 0] = src[0]; rgb_out[t * 3 + 1] = src[1]; rgb_out[t * 3 + 2] = src[2];
  }
  if (masks) mask_out[t] = masks[vi * npix + p];
}
}  // namespace lush

extern "C" int lush_ray_gen_gather(const float* poses, int n_views, int pose_stride, int pose_row_stride,
                                   const int64_t* view_idx, const int64_t* pix_idx, int H, int W, double fx, double fy,
                                   double cx, double cy, int64_t n, float* rays, const float* images, float* rgb_out,
                                   const uint8_t* masks, uint8_t* mask_out, int* err_flag, void* stream) {
  LUSH_REQUIRE(poses && n_views >= 1 && pose_row_stride >= 4 && pose_stride >= 2 * pose_row_stride + 4 && H > 0 &&
                   W > 0 && fx != 0.0 && fy != 0.0 && n >= 0 && (n == 0 || (view_idx && pix_idx && rays)),
               "ray_gen_gather: bad arguments");
  LUSH_REQUIRE((!images || rgb_out) && (!masks || mask_out), "ray_gen_gather: a table without its output buffer");
  if (n == 0) return 0;
  PinholeConst cam;
  cam.ax = (float)(0.5 - cx); cam.ay = (float)(0.5 - cy); cam.fx = (float)fx; cam.fy = (float)fy;
  int th = 256;
  ray_gen_gather_kernel<<<(unsigned)ceil_div(n, th), th, 0, (cudaStream_t)stream>>>(
      poses, n_views, pose_stride, pose_row_stride, view_idx, pix_idx, cam, W, (int64_t)H * W, n, rays, images, rgb_out,
      masks, mask_out, err_flag);
  LUSH_LAUNCH_CHECK("ray_gen_gather");
  return 0;
}

extern "C" int lush_ray_gen_prep(const float* c2w, int c2w_stride, int H, int W, double fx, double fy, double cx,
                                 double cy, float near, float far, int ndc, int64_t pix0, int64_t n,
                                 float* ray_batch, void* stream) {
  LUSH_REQUIRE(c2w && (ray_batch || n == 0) && c2w_stride >= 4 && H > 0 && W > 0 && fx != 0.0 && fy != 0.0,
               "ray_gen_prep: bad arguments");
  LUSH_REQUIRE(pix0 >= 0 && n >= 0 && pix0 + n <= (int64_t)H * W, "ray_gen_prep: pixel range outside the image");
  if (n == 0) return 0;
  PinholeConst cam;
  cam.ax = (float)(0.5 - cx); cam.ay = (float)(0.5 - cy); cam.fx = (float)fx; cam.fy = (float)fy;
  int th = 256;
  ray_gen_prep_kernel<<<(unsigned)ceil_div(n, th), th, 0, (cudaStream_t)stream>>>(
      c2w, c2w_stride, cam, W, pix0, n, ray_batch, make_ndc(H, W, fx), near, far, ndc);
  LUSH_LAUNCH_CHECK("ray_gen_prep");
  return 0;
}

extern "C" int lush_ray_prep_bwd(const float* rays, const float* g_ray_batch, float* g_rays,
                                 int64_t R, int H, int W, double focal, int ndc, void* stream) {
  LUSH_REQUIRE(rays && g_ray_batch && g_rays && R >= 0, "ray_prep_bwd: null pointer");
  if (R == 0) return 0;
  int th = 256;
  ray_prep_bwd_kernel<<<(unsigned)ceil_div(R, th), th, 0, (cudaStream_t)stream>>>(
      rays, g_ray_batch, g_rays, R, make_ndc(H, W, focal), ndc);
  LUSH_LAUNCH_CHECK("ray_prep_bwd");
  return 0;
}

extern "C" int lush_stratified_z(const float* ray_batch, int ray_stride, const float* t_vals,
                                 const float* t_rand, float* z_out, int64_t R, int S, int lindisp,
                                 void* stream) {
  LUSH_REQUIRE(ray_batch && t_vals && z_out && R >= 0 && S >= 1 && ray_stride >= 8,
               "stratified_z: bad arguments");
  if (R == 0) return 0;
  int th = 256;
  stratified_z_kernel<<<(unsigned)ceil_div(R * S, th), th, 0, (cudaStream_t)stream>>>(
      ray_batch, ray_stride, t_vals, t_rand, z_out, R, S, lindisp);
  LUSH_LAUNCH_CHECK("stratified_z");
  return 0;
}

extern "C" int lush_rbk_warp_fwd(const float* rays, const float* r, const float* v, float* new_rays,
                                 int64_t N, int M, void* stream) {
  LUSH_REQUIRE(rays && r && v && new_rays && N >= 0 && M >= 0, "rbk_warp_fwd: bad arguments");
  if (N == 0) return 0;
  int th = 128;
  rbk_warp_fwd_kernel<<<(unsigned)ceil_div(N * (M + 1), th), th, 0, (cudaStream_t)stream>>>(
      rays, r, v, new_rays, N, M);
  LUSH_LAUNCH_CHECK("rbk_warp_fwd");
  return 0;
}

extern "C" int lush_rbk_warp_bwd(const float* rays, const float* r, const float* v,
                                 const float* g_new_rays, float* g_rays, float* g_r, float* g_v,
                                 int64_t N, int M, void* stream) {
  LUSH_REQUIRE(rays && r && v && g_new_rays && g_rays && g_r && g_v && N >= 0 && M >= 0,
               "rbk_warp_bwd: bad arguments");
  if (N == 0) return 0;
  int th = 128;
  rbk_warp_bwd_kernel<<<(unsigned)ceil_div(N, th), th, 0, (cudaStream_t)stream>>>(
      rays, r, v, g_new_rays, g_rays, g_r, g_v, N, M);
  LUSH_LAUNCH_CHECK("rbk_warp_bwd");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Training loss 0.5*MSE + 0.5*L1 on the fine and the coarse blurred colours, value AND gradient in one
// launch (reference run_lushnerf.py:652-661: img2mse + img2mae on rgb_blur and rgb0_blur, 0.5 each).
// One CTA when the batch is small, so the sum order (and the loss bits) are fixed from run to run.
// ------------------------------------------------------------------------------------------------
__global__ void blur_loss_kernel(const float* __restrict__ blur, const float* __restrict__ blur0,
                                 const float* __restrict__ target, int64_t n, float inv_n,
                                 float* __restrict__ loss, float* __restrict__ g_blur,
                                 float* __restrict__ g_blur0) {
  __shared__ float s_part[32];
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float t = target[i];
    float d = blur[i] - t;
    acc += 0.5f * (d * d + fabsf(d));
    g_blur[i] = (d + 0.5f * (float)((d > 0.f) - (d < 0.f))) * inv_n;
    if (blur0) {
      float d0 = blur0[i] - t;
      acc += 0.5f * (d0 * d0 + fabsf(d0));
      g_blur0[i] = (d0 + 0.5f * (float)((d0 > 0.f) - (d0 < 0.f))) * inv_n;
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? s_part[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) {
      if (gridDim.x == 1) loss[0] = v * inv_n;
      else atomicAdd(loss, v * inv_n);
    }
  }
}

extern "C" int lush_blur_loss(const float* blur, const float* blur0, const float* target, int64_t n,
                              float* loss, float* g_blur, float* g_blur0, void* stream) {
  LUSH_REQUIRE(blur && target && loss && g_blur && n > 0, "blur_loss: bad arguments");
  LUSH_REQUIRE(!blur0 || g_blur0, "blur_loss: g_blur0 required with blur0");
  int th = 1024;
  unsigned blocks = n <= 32768 ? 1u : (unsigned)(ceil_div(n, (int64_t)th * 4) < 148 ? ceil_div(n, (int64_t)th * 4) : 148);
  if (blocks > 1) {
    cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), (cudaStream_t)stream);
    LUSH_REQUIRE(e == cudaSuccess, "blur_loss: memset failed: %s", cudaGetErrorString(e));
  }
  blur_loss_kernel<<<blocks, th, 0, (cudaStream_t)stream>>>(blur, blur0, target, n, 1.0f / (float)n, loss, g_blur,
                                                            g_blur0);
  LUSH_LAUNCH_CHECK("blur_loss");
  return 0;
}

extern "C" int lush_blend_tonemap_fwd(const float* rgb, const float* ccw, const float* noise_raw,
                                      float* out_blur, float* out_pure, float* noise_rgb, int64_t N,
                                      int K, int gamma, void* stream) {
  LUSH_REQUIRE(rgb && N >= 0 && K >= 1, "blend_tonemap_fwd: bad arguments");
  LUSH_REQUIRE(K == 1 || ccw, "blend_tonemap_fwd: ccw required when K > 1");
  if (N == 0) return 0;
  int th = 256;
  blend_tonemap_fwd_kernel<<<(unsigned)ceil_div(N * 3, th), th, 0, (cudaStream_t)stream>>>(
      rgb, ccw, noise_raw, out_blur, out_pure, noise_rgb, N, K, gamma);
  LUSH_LAUNCH_CHECK("blend_tonemap_fwd");
  return 0;
}

extern "C" int lush_blend_tonemap_bwd(const float* rgb, const float* ccw, const float* noise_raw,
                                      const float* g_blur, const float* g_pure,
                                      const float* g_noise_rgb, float* g_rgb, float* g_ccw,
                                      float* g_noise_raw, int64_t N, int K, int gamma, void* stream) {
  LUSH_REQUIRE(rgb && N >= 0 && K >= 1, "blend_tonemap_bwd: bad arguments");
  if (N == 0) return 0;
  int th = 128;
  blend_tonemap_bwd_kernel<<<(unsigned)ceil_div(N, th), th, 0, (cudaStream_t)stream>>>(
      rgb, ccw, noise_raw, g_blur, g_pure, g_noise_rgb, g_rgb, g_ccw, g_noise_raw, N, K, gamma);
  LUSH_LAUNCH_CHECK("blend_tonemap_bwd");
  return 0;
}
