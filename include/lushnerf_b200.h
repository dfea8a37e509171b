/*
 * lushnerf_b200.h -- C ABI of the B200-native LuSh-NeRF per-ray volumetric rendering hot path.
 *
 * The reference (quzefan/LuSh-NeRF) is pure PyTorch with no FFI of its own, so this header IS the
 * boundary a maintainer binds (ctypes stub in INTEGRATION.md).  Every entry point cites the reference
 * function it replaces as  file:line  relative to the reference root.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; all pointers are DEVICE pointers unless stated
 *   - all tensors are contiguous row-major fp32 unless stated; nullable pointers are marked
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream)
 *   - return value: 0 on success, negative LUSH_E* on argument errors, positive cudaError_t on a
 *     CUDA failure; never throws.  lush_last_error() gives a human-readable message (thread-local).
 *   - kernels never allocate: scratch comes in through `workspace` (size from the *_workspace_bytes
 *     query of the same op)
 */
#ifndef LUSHNERF_B200_H_
#define LUSHNERF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LUSH_ABI_VERSION 2

#define LUSH_OK 0
#define LUSH_EINVAL (-1)      /* bad argument / unsupported shape */
#define LUSH_EWORKSPACE (-2)  /* workspace too small */
#define LUSH_EUNSUPPORTED (-3)

/* Bits of the device-side error word (`flags` / `err_flag` arguments, int32, OR-ed atomically by the kernels and read
 * lazily by the host -- the replacement of the reference's ~20 isnan/isinf host syncs per render_rays call,
 * models/lushnerf.py:474-478, and of the index check nn.Embedding performs, :32,119):
 *   bits 0..9   NaN (even bit) / Inf (odd bit) in rgb_map, depth_map, acc_map, density_map, raw  (lush_composite_fwd)
 *   bit 10      an images_idx value outside [0, n_img)            (lush_rbk_motion_fwd)
 *   bit 11      a view index outside [0, n_views)                 (lush_ray_gen_gather)
 *   bit 12      a pixel index outside [0, H*W)                    (lush_ray_gen_gather)
 */
#define LUSH_FLAG_BAD_IMAGE_INDEX (1 << 10)
#define LUSH_FLAG_BAD_VIEW_INDEX (1 << 11)
#define LUSH_FLAG_BAD_PIXEL_INDEX (1 << 12)

int lush_abi_version(void);
const char* lush_last_error(void);
/* Number of kernels launched by this library since process start (bench.py: "gpu_launches"). */
int64_t lush_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Ray preparation.  Replaces the prologue shared by render_infer / render_train_scene /
 * render_train_noise (models/lushnerf.py:706-729, 772-795, 827-850) and ndc_rays
 * (utils/run_lushnerf_helpers.py:542-562):
 *   viewdirs = d/|d| (before NDC);  NDC warp with near plane 1 and focal K[0][0];
 *   ray_batch = [o(3) d(3) near far viewdirs(3)]  ->  [R,11]
 * rays is [R,3,2] (last axis: origin, direction).  ndc=0 skips the warp.
 * ------------------------------------------------------------------------------------------- */
int lush_ray_prep_fwd(const float* rays, float* ray_batch, int64_t R, int H, int W, double focal,
                      float near, float far, int ndc, void* stream);
/* g_ray_batch [R,11] -> g_rays [R,3,2] (written, not accumulated). */
int lush_ray_prep_bwd(const float* rays, const float* g_ray_batch, float* g_rays, int64_t R, int H,
                      int W, double focal, int ndc, void* stream);

/* Rays of pixels [pix0, pix0 + n) (row-major pixel index) of ONE camera pose, generated and prepared in one pass:
 * get_rays (reference utils/run_lushnerf_helpers.py:517-528) fused with the ray preparation above, replacing the
 * per-frame [H, W, 3, 2] ray image of render_path (models/lushnerf.py:881-882).  c2w: device pointer to the
 * [3, 4] (or [4, 4]) camera-to-world matrix, row stride c2w_stride floats.  Pixel centres at +0.5 like this
 * package's get_rays; the NDC constants use fx as the reference does (K[0][0]). */
int lush_ray_gen_prep(const float* c2w, int c2w_stride, int H, int W, double fx, double fy, double cx, double cy,
                      float near, float far, int ndc, int64_t pix0, int64_t n, float* ray_batch, void* stream);

/* A training batch staged on the device: for n (view, pixel) pairs, rays[t] = (o, d) of pixel pix_idx[t] (row-major
 * index in [0, H*W)) seen from poses[view_idx[t]], and -- optionally -- the pixel's target colour and frequency-mask
 * bit gathered from device-resident tables.  Replaces the pre-expanded [N, ro+rd+rgb, 3] ray table, its per-epoch
 * host-side shuffles and the per-iteration slicing (reference run_lushnerf.py:538-616 with get_rays_np,
 * helpers:531-539): only the two index vectors travel per step.
 *   poses   [n_views, rows >= 3, row_stride >= 4] fp32, view v at poses + v*pose_stride, row r at + r*pose_row_stride
 *           ([V,3,4]: 12 / 4;  [V,4,4]: 16 / 4;  the LLFF [V,3,5] block: 15 / 5)
 *   rays    [n, 3, 2] fp32 as NeRFAll.forward takes them
 *   images  nullable [n_views, H*W, 3] fp32 -> rgb_out [n, 3];   masks nullable [n_views, H*W] uint8 -> mask_out [n]
 *   err_flag nullable device int32: an index outside its range is CLAMPED (no out-of-bounds read, a defined ray) and
 *           LUSH_FLAG_BAD_VIEW_INDEX / LUSH_FLAG_BAD_PIXEL_INDEX is OR-ed into it */
int lush_ray_gen_gather(const float* poses, int n_views, int pose_stride, int pose_row_stride, const int64_t* view_idx,
                        const int64_t* pix_idx, int H, int W, double fx, double fy, double cx, double cy, int64_t n,
                        float* rays, const float* images, float* rgb_out, const uint8_t* masks, uint8_t* mask_out,
                        int* err_flag, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Coarse sample placement.  Replaces models/lushnerf.py:389-412 (render_rays) / :501-523:
 *   t = linspace(0,1,S) (passed in as t_vals[S], computed by the host exactly as the reference
 *   does), z = near(1-t)+far t (or the lindisp form), optional stratified jitter with t_rand[R,S]
 *   (nullable: no perturbation).  z_out [R,S].
 * ------------------------------------------------------------------------------------------- */
int lush_stratified_z(const float* ray_batch, int ray_stride, const float* t_vals,
                      const float* t_rand, float* z_out, int64_t R, int S, int lindisp, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Alpha compositing.  Replaces NeRFAll.raw2outputs (models/lushnerf.py:296-352).
 *   raw [R,S,4], z [R,S], rays_d: R rows of 3 floats with row stride `d_stride` floats,
 *   sigma_noise [R,S-1] nullable (already multiplied by raw_noise_std),
 *   rm_near_thresh: density_i *= (z_{i+1} > thresh) when >= 0 (eval branch, :331-335; pass -1 in
 *   training), white_bkgd: rgb += 1-acc (:349-350).
 *   outputs: rgb_map [R,3], density [R,S-1], acc [R], weights [R,S], depth [R]
 *   (density / weights nullable when the caller does not need them).
 *   flags (nullable int32[1]): bit0 |= any NaN, bit1 |= any Inf in the outputs (replaces the
 *   isnan/isinf host syncs at :474-478).
 * ------------------------------------------------------------------------------------------- */
int lush_composite_fwd(const float* raw, const float* z, const float* rays_d, int d_stride,
                       const float* sigma_noise, float rm_near_thresh, int white_bkgd,
                       float* rgb_map, float* density, float* acc, float* weights, float* depth,
                       int* flags, int64_t R, int S, void* stream);
/* Backward of the above (closed form).  Upstream grads (all nullable = zero):
 *   g_rgb [R,3], g_density [R,S-1], g_acc [R], g_weights [R,S], g_depth [R].
 * Outputs: g_raw [R,S,4] (written), g_rays_d [R,3] nullable (written; dense stride 3). */
int lush_composite_bwd(const float* raw, const float* z, const float* rays_d, int d_stride,
                       const float* sigma_noise, float rm_near_thresh, int white_bkgd,
                       const float* g_rgb, const float* g_density, const float* g_acc,
                       const float* g_weights, const float* g_depth, float* g_raw, float* g_rays_d,
                       int64_t R, int S, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Hierarchical sampling.  Replaces sample_pdf (utils/run_lushnerf_helpers.py:566-609).
 *   bins [R,nb], weights [R,nb-1]  (nb = S-1 at the render_rays call site, :435-436)
 *   u: det -> u is [N] shared by all rays (u_stride = 0); random -> [R,N] (u_stride = N)
 *   cdf_in (nullable) [R,nb]: STRICT mode, the caller supplies the reference-computed cdf and the
 *     kernel only searches/gathers/lerps.  When null the cdf is built in-kernel with the association
 *     order documented in DESIGN.md (bit-exact against oracle.sample_pdf(order="kernel")).
 *   samples [R,N];  inds_out (nullable) int64 [R,N] = searchsorted(cdf, u, right=True).
 * ------------------------------------------------------------------------------------------- */
int lush_sample_pdf(const float* bins, const float* weights, const float* cdf_in, const float* u,
                    int64_t u_stride, float* samples, int64_t* inds_out, int64_t R, int nb, int N,
                    void* stream);
/* Fused fine-sample placement: z_mid + sample_pdf(weights[:,1:-1]) + sort(cat(z, z_samples))
 * (models/lushnerf.py:435-440 / :544-549) + z_std (:465).  z [R,S], weights [R,S] (full weights of the
 * coarse pass; nullable when cdf_in is given).  cdf_in (nullable) [R,S-1]: STRICT mode as in lush_sample_pdf -- the
 * cdf of weights[:,1:-1] computed by the caller with the reference's own ops (helpers:568-571).
 * Outputs z_fine [R,S+N] sorted ascending, z_samples [R,N], z_std [R] nullable.
 * Precondition: every row of z is ascending and free of NaNs (the coarse samples of render_rays are); the merge is
 * by rank, not a full sort.  If it is violated the unfilled slots of z_fine hold NaN (never stale memory). */
int lush_sample_merge(const float* z, const float* weights, const float* cdf_in, const float* u, int64_t u_stride,
                      float* z_fine, float* z_samples, float* z_std, int64_t R, int S, int N,
                      void* stream);

/* ---------------------------------------------------------------------------------------------
 * Encoder + MLP.  Replaces Embedder.forward (helpers:311-344), NeRF.forward (helpers:394-423),
 * NeRF_Noise.forward (helpers:483-512), NeRFAll.mlpforward / mlpforward_noise
 * (models/lushnerf.py:234-293).
 *
 * lush_mlp_desc describes the network; `params` is ONE flat fp32 buffer holding the tensors in
 * state_dict order:  pts_linears.{0..D-1}.{weight,bias}, views_linears.0.{weight,bias},
 * feature_linear.{weight,bias}, alpha_linear.{weight,bias}, rgb_linear.{weight,bias}
 * (each weight [out,in] row-major, as nn.Linear stores it).
 * ------------------------------------------------------------------------------------------- */
typedef struct lush_mlp_desc {
  int32_t D;           /* number of pts_linears */
  int32_t W;           /* hidden width */
  int32_t L_pts;       /* multires (10) -> input_ch = 3 + 6*L_pts */
  int32_t L_views;     /* multires_views (4) -> input_ch_views = 3 + 6*L_views */
  int32_t skip;        /* skip layer index (4); -1 = none */
  int32_t out_mode;    /* 0: scene MLP, out = [rgb(3), sigma(1)];  1: noise MLP, out = rgb(3) */
} lush_mlp_desc;

/* Number of fp32 elements in the flat parameter buffer for `desc`. */
int64_t lush_mlp_param_count(const lush_mlp_desc* desc);
/* Capability bits of `impl` (0 = fp32 SIMT, 1 = tcgen05 fp16-operand) for this topology:
 * bit 0 = forward / inference, bit 1 = training (forward with stash + backward).  0 = unsupported. */
int lush_mlp_impl_supported(const lush_mlp_desc* desc, int impl);

/* Sample-point generation is fused in: for ray r and sample s the network input is
 *   x = o_r + d_r * z[r,s]   (models/lushnerf.py:414 / :441), viewdir v_r shared by the ray.
 * ray_batch [R, ray_stride] with o at col 0, d at col 3, viewdirs at col 8.  z [R,S].
 * For the noise MLP (out_mode 1) the reference evaluates ONE point per ray, sample index 16 of the
 * un-jittered coarse z (models/lushnerf.py:271): pass S=1 and z = that column.
 * out: [R*S, 4] (mode 0) or [R*S, 3] (mode 1).
 * impl: 0 = fp32 SIMT (exact-fp32 parity mode), 1 = fp16-operand/fp32-accumulate tcgen05 path.
 * `stash` (nullable) receives the activations needed by lush_mlp_bwd (size from
 * lush_mlp_stash_bytes); pass null for inference. */
size_t lush_mlp_workspace_bytes(const lush_mlp_desc* desc, int64_t R, int S, int impl, int training);
size_t lush_mlp_stash_bytes(const lush_mlp_desc* desc, int64_t R, int S, int impl);
int lush_mlp_fwd(const lush_mlp_desc* desc, const float* params, const float* ray_batch,
                 int ray_stride, const float* z, float* out, void* stash, void* workspace,
                 size_t workspace_bytes, int64_t R, int S, int impl, void* stream);
/* Backward.  g_out like out.  g_params: flat fp32 buffer like params, ACCUMULATED into (caller
 * zeroes).  g_ray_batch (nullable) [R,11]: written with d/d(o), d/d(d), 0, 0, d/d(viewdirs). */
int lush_mlp_bwd(const lush_mlp_desc* desc, const float* params, const float* ray_batch,
                 int ray_stride, const float* z, const float* g_out, const void* stash,
                 float* g_params, float* g_ray_batch, void* workspace, size_t workspace_bytes,
                 int64_t R, int S, int impl, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Rigid-warp blur kernel.  Replaces Rigid_Blurring_Kernel.rbk_warp (models/lushnerf.py:75-98) and
 * SE3Field.warp / Rigid_body.exp_se3 (utils/rigid_warping.py:20-140).
 *   rays [N,3,2]; r, v [N,3,M] (already multiplied by rv_window, as at :144-145);
 *   new_rays [N, M+1, 3, 2] with slot 0 = the original ray (use_origin, :83-84).
 * ------------------------------------------------------------------------------------------- */
int lush_rbk_warp_fwd(const float* rays, const float* r, const float* v, float* new_rays, int64_t N,
                      int M, void* stream);
int lush_rbk_warp_bwd(const float* rays, const float* r, const float* v, const float* g_new_rays,
                      float* g_rays, float* g_r, float* g_v, int64_t N, int M, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K-ray blend + noise add + tone map.  Replaces rbk_weighted_sum (models/lushnerf.py:100-116) for
 * rgb, the 0.1*sigmoid(noise) add (:649) and ToneMapping('gamma') (helpers:164-174):
 *   blended[n] = sum_k ccw[n,k] * rgb[n*K+k];  noise_rgb = 0.1*sigmoid(noise_raw) (nullable)
 *   out_blur = tm(blended + noise_rgb), out_pure = tm(blended);  tm(x) = x^(1/2.2) if gamma else x
 * K=1 with ccw null degenerates to tone mapping only.
 * ------------------------------------------------------------------------------------------- */
int lush_blend_tonemap_fwd(const float* rgb, const float* ccw, const float* noise_raw,
                           float* out_blur, float* out_pure, float* noise_rgb, int64_t N, int K,
                           int gamma, void* stream);
int lush_blend_tonemap_bwd(const float* rgb, const float* ccw, const float* noise_raw,
                           const float* g_blur, const float* g_pure, const float* g_noise_rgb,
                           float* g_rgb, float* g_ccw, float* g_noise_raw, int64_t N, int K,
                           int gamma, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Motion-parameter network of the rigid blurring kernel.
 * Replaces Rigid_Blurring_Kernel.forward up to the warp call (reference models/lushnerf.py:118-147) and the
 * autograd graph of those lines: view embedding -> trunk of ReLU layers (optional skip concat [e, h]) ->
 * r / v / w branches -> r = lin*rv_window, v = lin*rv_window, w = sigmoid(lin), ccw = w / (sum(w) + 1e-10).
 * The result depends on images_idx alone, so the network is evaluated once per row of the embedding table and
 * every ray receives the row of its image.
 *
 * Activation row (row_width floats per image): columns [0, embed_dim) hold the embedding, every layer writes
 * its out_dim outputs at out_off and reads in0 (and optionally in1, the concat partner) at columns that precede
 * out_off.  Weights are torch nn.Linear tensors: w [out_dim, in0_dim + in1_dim] row-major, b [out_dim].
 * r_off / v_off / w_off are the out_off of the three head layers (act = 0, widths 3M, 3M, M+1).
 *   fwd: rows [n_img, row_width] receives the activation rows (needed by bwd); r, v [N, 3M]; ccw [N, M+1].
 *   bwd: g_r / g_v / g_ccw may be null (no gradient); gw / gb of every layer are ACCUMULATED into (zero them
 *        first); g_embed [n_img, embed_dim] is written (may be null).
 * An images_idx value outside [0, n_img) (nn.Embedding raises on it): fwd gives that ray the identity kernel
 * (r = v = 0, ccw = (1, 0, ..., 0)) and ORs LUSH_FLAG_BAD_IMAGE_INDEX into *err_flag (nullable device int32); bwd
 * ignores the ray.
 * ------------------------------------------------------------------------------------------- */
#define LUSH_MOTION_MAX_LAYERS 24
typedef struct lush_dense_layer {
  const float* w;
  const float* b;
  float* gw;
  float* gb;
  int32_t in0_off, in0_dim, in1_off, in1_dim, out_off, out_dim, act; /* act: 0 = identity, 1 = ReLU */
} lush_dense_layer;
typedef struct lush_motion_net {
  int32_t n_layers, row_width, embed_dim, num_motion;
  int32_t r_off, v_off, w_off;
  float rv_window;
  lush_dense_layer layers[LUSH_MOTION_MAX_LAYERS];
} lush_motion_net;
int lush_rbk_motion_fwd(const lush_motion_net* net, const float* embed, int n_img, const int64_t* images_idx,
                        int64_t N, float* rows, float* r, float* v, float* ccw, int* err_flag, void* stream);
int lush_rbk_motion_bwd(const lush_motion_net* net, int n_img, const int64_t* images_idx, int64_t N,
                        const float* rows, const float* g_r, const float* g_v, const float* g_ccw,
                        float* g_embed, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training loss of the blurred colours, value and gradient in one launch.
 * Replaces reference run_lushnerf.py:652-661 (img2mse / img2mae of rgb_blur and rgb0_blur against the
 * target pixels, weighted 0.5 each) together with the autograd nodes those lines create.
 *   loss[0]   = 0.5*mean(d^2) + 0.5*mean|d| + 0.5*mean(d0^2) + 0.5*mean|d0|,  d = blur - target, d0 = blur0 - target
 *   g_blur[i] = (d_i + 0.5*sign(d_i)) / n      (same for g_blur0);  blur0 / g_blur0 may be null (single term pair)
 * n = number of floats (3 * pixels).  n <= 32768 runs as one CTA: bit-reproducible sum order.
 * ------------------------------------------------------------------------------------------- */
int lush_blur_loss(const float* blur, const float* blur0, const float* target, int64_t n, float* loss,
                   float* g_blur, float* g_blur0, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-tensor Adam step.  Replaces the reference's optimizer = torch.optim.Adam(params=grad_vars, lr=lrate,
 * betas=(0.9, 0.999)) and its optimizer.step() (run_lushnerf.py:359-371, :675-677): no weight decay, no amsgrad.
 *   exp_avg += (1-beta1)*(grad-exp_avg);  exp_avg_sq = beta2*exp_avg_sq + (1-beta2)*grad^2
 *   param  -= lr/(1-beta1^t) * exp_avg / (sqrt(exp_avg_sq)/sqrt(1-beta2^t) + eps),  t = steps[step_index] + 1
 * `segs` is a HOST array (one entry per parameter tensor that has a gradient); lr[] (one learning rate per
 * parameter group), steps[] (updates received so far, per tensor) and done_counter (one zero-initialised
 * word) live on the DEVICE, so a captured CUDA graph follows learning-rate changes and keeps counting.
 * The kernel increments steps[step_index] of every segment once.  Segments must not overlap.
 * ------------------------------------------------------------------------------------------- */
#define LUSH_ADAM_MAX_SEGS 64   /* tensors per launch; longer lists are split into several launches */
typedef struct lush_adam_seg {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t numel;
  int32_t step_index;
  int32_t lr_index;
} lush_adam_seg;
int lush_adam_step(const lush_adam_seg* segs, int nseg, const float* lr, float* steps,
                   unsigned int* done_counter, double beta1, double beta2, double eps, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Diagnostics (not on the product path): one-CTA tcgen05 GEMM D[128,N] = A[128,K] . B[N,K]^T over fp16
 * "tile-canonical" operand blobs, used by the GPU tests to pin the UMMA descriptor conventions that the
 * fused MLP kernels rely on.  a_mode / b_mode: 0 = K-major view, 1 = MN-major view of the blob.
 * ------------------------------------------------------------------------------------------- */
int lush_tc_selftest(const void* a_blob, const void* b_blob, float* d_out, int N, int K, int a_mode,
                     int b_mode, int swap_lbo_sbo, void* stream);

/* Tensor-pipe micro-benchmark (tools/bench_umma.py; not on the product path): `iters` back-to-back
 * tcgen05.mma M=128 K=16 from one thread; out_cycles[0] = issue cycles, out_cycles[1] = cycles to completion.
 * flags (per block of 8 MMAs): 1 = tcgen05.commit, 2 = tcgen05.fence::after_thread_sync, 4 = mbarrier wait. */
int lush_tc_microbench(long long* out_cycles, int N, int nacc, int a_tmem, int iters, int flags, void* stream);
/* Diagnostics: CTA-pair UMMA (cta_group::2).  D[256 x N] = A[256 x K] . B[N x K]^T on a cluster of two CTAs; a_blob =
 * two consecutive K-major canonical [128 x K] blobs (rows 0..127, 128..255), b_blob = two consecutive K-major
 * canonical [N/2 x K] blobs; d_out [256, N] fp32 row-major.  Pins the conventions a 2-CTA chain kernel needs. */
int lush_tc_selftest_pair(const void* a_blob, const void* b_blob, float* d_out, int N, int K, void* stream);

/* Diagnostics: every CTA streams the same nblocks x block_bytes image from L2 through a 3-stage shared-memory ring
 * `iters` times, unicast (cluster = 1) or with each CTA of a cluster fetching 1/cluster of a block and multicasting
 * it (cluster = 2, 4).  out_cycles[grid] = cycles per CTA.  Measures the weight-delivery rate of the chain kernels. */
int lush_tc_bcast_bench(const void* src, int nblocks, int block_bytes, int iters, int cluster, int grid,
                        long long* out_cycles, void* stream);

/* Debug (library built with -DLUSH_CHAIN_TRACE): copy the forward chain kernel's timeline trace (LUSH_CHAIN_DEBUG=2; clock64 stamps of CTA 0, third tile,
 * slot = step*16 + event) to host memory. */
int lush_debug_read_chain_trace(long long* host_out, int n);
int lush_debug_read_chain_trace_bwd(long long* host_out, int n);   /* same for the dgrad chain kernel */

#ifdef __cplusplus
}
#endif
#endif /* LUSHNERF_B200_H_ */
