// Multi-tensor Adam: every parameter tensor of the training step updated by ONE launch per 64 tensors.
// Replaces the reference's torch.optim.Adam(params=grad_vars, lr, betas=(0.9, 0.999)).step()
// (run_lushnerf.py:359-371 construction, :675-677 zero_grad / backward / step).
//
// torch's own fused multi-tensor Adam hands one 65536-element chunk to one CTA: the 256x256 layers of the two
// scene networks become ~20 CTAs on a 148-SM part and the update of 1.3 M parameters takes ~190 us.  Here the
// tensors form one virtual index space that a full grid walks with coalesced scalar accesses (~36 MB moved).
#include "common.cuh"

namespace lush {

struct AdamSegs {
  float* p[LUSH_ADAM_MAX_SEGS];
  const float* g[LUSH_ADAM_MAX_SEGS];
  float* m[LUSH_ADAM_MAX_SEGS];
  float* v[LUSH_ADAM_MAX_SEGS];
  long long end[LUSH_ADAM_MAX_SEGS];   // exclusive end of the segment in the virtual index space
  int step_idx[LUSH_ADAM_MAX_SEGS];    // slot of this tensor in steps[]
  int lr_idx[LUSH_ADAM_MAX_SEGS];      // slot of this tensor's group in lr[]
  int n;
};

// steps[] holds the number of updates each tensor has RECEIVED so far (fp32 like torch's capturable step
// tensors); the last CTA to finish bumps the slots of this launch, so the whole update is one graph node.
__global__ void __launch_bounds__(256)
adam_kernel(const __grid_constant__ AdamSegs S, const float* __restrict__ lr, float* __restrict__ steps,
            unsigned int* __restrict__ done_counter, double beta1, double beta2, double eps) {
  __shared__ long long s_end[LUSH_ADAM_MAX_SEGS];
  __shared__ float s_step_size[LUSH_ADAM_MAX_SEGS], s_bc2_sqrt[LUSH_ADAM_MAX_SEGS];
  __shared__ bool s_last;
  for (int s = threadIdx.x; s < S.n; s += blockDim.x) {
    s_end[s] = S.end[s];
    double t = (double)steps[S.step_idx[s]] + 1.0;
    double bc1 = 1.0 - pow(beta1, t);
    double bc2 = 1.0 - pow(beta2, t);
    s_step_size[s] = (float)((double)lr[S.lr_idx[s]] / bc1);
    s_bc2_sqrt[s] = (float)sqrt(bc2);
  }
  __syncthreads();
  const long long total = s_end[S.n - 1];
  const float w1 = (float)(1.0 - beta1), feps = (float)eps;
  int s = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    while (i >= s_end[s]) ++s;
    long long k = i - (s ? s_end[s - 1] : 0);
    float g = S.g[s][k];
    float m = S.m[s][k], v = S.v[s][k];
    // same roundings as torch's fused kernel: betas are doubles there, so the second moment is formed in fp64
    m = m + w1 * (g - m);                            // exp_avg.lerp_(grad, 1 - beta1)
    v = (float)(beta2 * (double)v + ((1.0 - beta2) * (double)g) * (double)g);
    float denom = sqrtf(v) / s_bc2_sqrt[s] + feps;
    S.m[s][k] = m;
    S.v[s][k] = v;
    S.p[s][k] -= s_step_size[s] * (m / denom);
  }
  // all CTAs have read steps[] before any of them reaches the counter: the last one to arrive advances the counters
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (s_last) {
    for (int q = threadIdx.x; q < S.n; q += blockDim.x) steps[S.step_idx[q]] += 1.f;
    if (threadIdx.x == 0) *done_counter = 0u;
  }
}

}  // namespace lush

using namespace lush;

extern "C" int lush_adam_step(const lush_adam_seg* segs, int nseg, const float* lr, float* steps,
                              unsigned int* done_counter, double beta1, double beta2, double eps, void* stream) {
  LUSH_REQUIRE(nseg >= 0 && (nseg == 0 || segs) && lr && steps && done_counter, "adam_step: bad arguments");
  for (int base = 0; base < nseg; base += LUSH_ADAM_MAX_SEGS) {
    AdamSegs S;
    S.n = nseg - base < LUSH_ADAM_MAX_SEGS ? nseg - base : LUSH_ADAM_MAX_SEGS;
    long long total = 0;
    for (int i = 0; i < S.n; ++i) {
      const lush_adam_seg& a = segs[base + i];
      LUSH_REQUIRE(a.param && a.grad && a.exp_avg && a.exp_avg_sq && a.numel > 0 && a.step_index >= 0 &&
                       a.lr_index >= 0,
                   "adam_step: segment %d is incomplete", base + i);
      S.p[i] = a.param; S.g[i] = a.grad; S.m[i] = a.exp_avg; S.v[i] = a.exp_avg_sq;
      total += a.numel;
      S.end[i] = total;
      S.step_idx[i] = a.step_index;
      S.lr_idx[i] = a.lr_index;
    }
    long long want = ceil_div(total, (long long)256 * 4);
    unsigned blocks = (unsigned)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
    adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(S, lr, steps, done_counter, beta1, beta2, eps);
    LUSH_LAUNCH_CHECK("adam_step");
  }
  return 0;
}
