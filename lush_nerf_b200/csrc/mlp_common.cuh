// Parameter-buffer layout shared by the SIMT (exact fp32) and tcgen05 (fp16 operand) MLP paths.
// Flat fp32 buffer in state_dict order (reference utils/run_lushnerf_helpers.py:379-390):
//   pts_linears.{i}.weight [W, in_i], .bias [W]   (in_0 = in_ch; in_{skip+1} = W + in_ch; else W)
//   views_linears.0.weight [W/2, W + in_ch_v], .bias [W/2]
//   feature_linear.weight [W, W], .bias [W]
//   alpha_linear.weight [1, W], .bias [1]
//   rgb_linear.weight [3, W/2], .bias [3]
#pragma once
#include "common.cuh"

namespace lush {

constexpr int kMaxDepth = 16;

struct MlpLayout {
  int D, W, in_ch, in_ch_v, skip, out_mode;
  int pts_in[kMaxDepth];       // fan-in of pts_linears[i]
  int64_t pts_w[kMaxDepth], pts_b[kMaxDepth];
  int64_t views_w, views_b, feat_w, feat_b, alpha_w, alpha_b, rgb_w, rgb_b;
  int64_t total;
  int out_ch;                  // 4 (scene) or 3 (noise)
};

inline bool make_layout(const lush_mlp_desc* d, MlpLayout& L) {
  if (!d || d->D < 1 || d->D > kMaxDepth || d->W < 2 || (d->W & 1) || d->L_pts < 0 || d->L_views < 0)
    return false;
  L.D = d->D; L.W = d->W; L.skip = d->skip; L.out_mode = d->out_mode;
  L.in_ch = 3 + 6 * d->L_pts;
  L.in_ch_v = 3 + 6 * d->L_views;
  int64_t off = 0;
  for (int i = 0; i < L.D; ++i) {
    // reference: Linear(W + input_ch, W) for layer i+1 when i in skips (helpers:380)
    L.pts_in[i] = (i == 0) ? L.in_ch : ((d->skip >= 0 && i == d->skip + 1) ? L.W + L.in_ch : L.W);
    L.pts_w[i] = off; off += (int64_t)L.W * L.pts_in[i];
    L.pts_b[i] = off; off += L.W;
  }
  L.views_w = off; off += (int64_t)(L.W / 2) * (L.W + L.in_ch_v);
  L.views_b = off; off += L.W / 2;
  L.feat_w = off; off += (int64_t)L.W * L.W;
  L.feat_b = off; off += L.W;
  L.alpha_w = off; off += L.W;
  L.alpha_b = off; off += 1;
  L.rgb_w = off; off += 3 * (L.W / 2);
  L.rgb_b = off; off += 3;
  L.total = off;
  L.out_ch = (d->out_mode == 0) ? 4 : 3;
  return true;
}

// input width of the tensor feeding pts_linears[D-1]'s consumers (alpha / feature heads):
// if the LAST layer index is in skips the reference concatenates again (helpers:400-401).
inline int trunk_out_width(const MlpLayout& L) {
  return (L.skip == L.D - 1) ? L.W + L.in_ch : L.W;
}

}  // namespace lush
