// Hierarchical inverse-CDF resampling as a warp-per-ray search kernel, optionally fused with the
// z_mid construction, the coarse/fine z merge (sort) and z_std.
// Replaces sample_pdf (reference utils/run_lushnerf_helpers.py:566-609) and
// models/lushnerf.py:435-440,465.
//
// Bit-exactness: this file is compiled with --fmad=false and every step is a single IEEE fp32 op, so
// given identical cdf/u the indices (searchsorted right=True) and samples equal the torch result
// bit for bit.  The in-kernel cdf uses a FIXED association order ("kernel order", see
// warp_cdf_build) that oracle/lush_oracle.py::sample_pdf(order="kernel") mirrors exactly; strict mode
// (cdf_in != null) consumes a caller-provided cdf instead.
#include "common.cuh"

namespace lush {

constexpr int kSamplerWarps = 4;

// Build cdf[0..n] (cdf[0] = 0) from n weights held in shared memory `w` (raw weights, the 1e-5 is
// added here).  Kernel order:
//   E = ceil(n/32); lane L owns elements [L*E, (L+1)*E)
//   lane partial = left-to-right fp32 sum;  total = xor-butterfly (16,8,4,2,1) of the partials
//   pdf = w' / total;  per-lane left-to-right running sum c;  lane totals scanned Kogge-Stone
//   (1,2,4,8,16, v = v + shfl_up(v,o));  cdf[j+1] = offset_L + c
__device__ __forceinline__ void warp_cdf_build(const float* __restrict__ w, float* __restrict__ cdf,
                                               int n, int lane) {
  const int E = (n + 31) / 32;
  const int j0 = lane * E;
  float part = 0.f;
  for (int e = 0; e < E; ++e) {
    int j = j0 + e;
    if (j < n) part = part + (w[j] + 1e-5f);
  }
  float total = part;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) total = total + __shfl_xor_sync(0xffffffffu, total, o);
  float run = 0.f;
  for (int e = 0; e < E; ++e) {
    int j = j0 + e;
    if (j < n) run = run + (w[j] + 1e-5f) / total;
  }
  float v = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v = v + t;
  }
  float off = __shfl_up_sync(0xffffffffu, v, 1);
  if (lane == 0) off = 0.f;
  run = 0.f;
  for (int e = 0; e < E; ++e) {
    int j = j0 + e;
    if (j < n) {
      run = run + (w[j] + 1e-5f) / total;
      cdf[j + 1] = off + run;
    }
  }
  if (lane == 0) cdf[0] = 0.f;
}

// searchsorted(cdf[0..nb), u, right=True): number of entries <= u
__device__ __forceinline__ int upper_bound(const float* __restrict__ cdf, int nb, float u) {
  int lo = 0, hi = nb;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ float invert_cdf(const float* __restrict__ cdf,
                                            const float* __restrict__ bins, int nb, float u,
                                            int& inds) {
  inds = upper_bound(cdf, nb, u);
  int below = max(0, inds - 1);
  int above = min(nb - 1, inds);
  float c0 = cdf[below], c1 = cdf[above];
  float b0 = bins[below], b1 = bins[above];
  float denom = c1 - c0;
  if (denom < 1e-5f) denom = 1.f;
  float t = (u - c0) / denom;
  return b0 + t * (b1 - b0);
}

// smem per warp: cdf[nb] + bins[nb] + w[nb-1]  (rounded up)
__global__ void __launch_bounds__(kSamplerWarps * 32)
sample_pdf_kernel(const float* __restrict__ bins, const float* __restrict__ weights,
                  const float* __restrict__ cdf_in, const float* __restrict__ u, int64_t u_stride,
                  float* __restrict__ samples, int64_t* __restrict__ inds_out, int64_t R, int nb,
                  int N, int per_warp) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * kSamplerWarps + wid;
  if (r >= R) return;
  float* s_cdf = smem + wid * per_warp;
  float* s_bins = s_cdf + nb;
  float* s_w = s_bins + nb;
  for (int j = lane; j < nb; j += 32) s_bins[j] = bins[r * nb + j];
  if (cdf_in != nullptr) {
    for (int j = lane; j < nb; j += 32) s_cdf[j] = cdf_in[r * nb + j];
    __syncwarp();
  } else {
    for (int j = lane; j < nb - 1; j += 32) s_w[j] = weights[r * (nb - 1) + j];
    __syncwarp();
    warp_cdf_build(s_w, s_cdf, nb - 1, lane);
    __syncwarp();
  }
  for (int k = lane; k < N; k += 32) {
    float uu = u[r * u_stride + k];
    int inds;
    float s = invert_cdf(s_cdf, s_bins, nb, uu, inds);
    samples[r * N + k] = s;
    if (inds_out) inds_out[r * N + k] = inds;
  }
}

// ascending bitonic sort of 32*KPL keys held in shared memory by one warp: lane l owns keys [l*KPL, (l+1)*KPL)
// in registers; stages with partner distance >= KPL exchange through shuffles, smaller ones stay in the lane
template <int KPL>
__device__ __forceinline__ void warp_bitonic_sort_regs(float* __restrict__ keys, int lane) {
  float v[KPL];
#pragma unroll
  for (int q = 0; q < KPL; ++q) v[q] = keys[lane * KPL + q];
#pragma unroll
  for (int k = 2; k <= 32 * KPL; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= KPL) {
        const int lj = j / KPL;                                // partner lane distance
        const bool lower = (lane & lj) == 0;                   // this lane holds the lower index of each pair
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
          const int e = lane * KPL + q;
          const bool up = (e & k) == 0;
          const float o = __shfl_xor_sync(0xffffffffu, v[q], lj);
          const float mn = fminf(v[q], o), mx = fmaxf(v[q], o);
          v[q] = (lower == up) ? mn : mx;
        }
      } else {
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
          if ((q & j) == 0) {
            const int e = lane * KPL + q;
            const bool up = (e & k) == 0;
            const float a = v[q], b = v[q | j];
            const float mn = fminf(a, b), mx = fmaxf(a, b);
            v[q] = up ? mn : mx;
            v[q | j] = up ? mx : mn;
          }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < KPL; ++q) keys[lane * KPL + q] = v[q];
  __syncwarp();
}

// Fused: z_mid, cdf from weights[1:-1], N samples, sort(cat(z, samples)), z_std.
// The coarse z of a ray are already ascending (stratified bins), so only the N new samples are sorted (bitonic,
// PN = next pow2 >= N keys) and the two ascending lists are MERGED by rank: z[i] lands at i + #{samples < z[i]},
// sample k at k + #{z <= sample_k} (binary searches in shared memory) -- the same values torch.sort produces.
// smem per warp: cdf[S-1] + bins[S-1] + w[S-1] + z[S] + samples[PN] + out[S+N]
__global__ void __launch_bounds__(kSamplerWarps * 32)
sample_merge_kernel(const float* __restrict__ z, const float* __restrict__ weights,
                    const float* __restrict__ cdf_in, const float* __restrict__ u, int64_t u_stride,
                    float* __restrict__ z_fine,
                    float* __restrict__ z_samples, float* __restrict__ z_std, int64_t R, int S, int N,
                    int PN, int per_warp) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * kSamplerWarps + wid;
  if (r >= R) return;
  const int nb = S - 1;
  float* s_cdf = smem + wid * per_warp;
  float* s_bins = s_cdf + nb;
  float* s_w = s_bins + nb;
  float* s_z = s_w + nb;       // (nb-1 used) keep simple alignment
  float* s_smp = s_z + S;
  float* s_out = s_smp + PN;
  for (int j = lane; j < S; j += 32) s_z[j] = z[r * S + j];
  if (cdf_in != nullptr) {      // strict mode: the caller's cdf [R, S-1] (e.g. the reference's torch.cumsum on this device)
    for (int j = lane; j < nb; j += 32) s_cdf[j] = cdf_in[r * nb + j];
  } else {
    for (int j = lane; j < S - 2; j += 32) s_w[j] = weights[r * S + j + 1];
  }
  __syncwarp();
  for (int j = lane; j < nb; j += 32) s_bins[j] = 0.5f * (s_z[j + 1] + s_z[j]);
  if (cdf_in == nullptr) warp_cdf_build(s_w, s_cdf, nb - 1, lane);
  __syncwarp();
  float sum = 0.f;
  for (int k = lane; k < N; k += 32) {
    float uu = u[r * u_stride + k];
    int inds;
    float s = invert_cdf(s_cdf, s_bins, nb, uu, inds);
    s_smp[k] = s;
    z_samples[r * N + k] = s;
    sum += s;
  }
  for (int j = N + lane; j < PN; j += 32) s_smp[j] = INFINITY;
  if (z_std != nullptr) {  // population std (torch.std(unbiased=False))
    sum = warp_sum(sum);
    float mean = sum / (float)N, sq = 0.f;
    __syncwarp();
    for (int k = lane; k < N; k += 32) {
      float d = s_smp[k] - mean;
      sq += d * d;
    }
    sq = warp_sum(sq);
    if (lane == 0) z_std[r] = sqrtf(sq / (float)N);
  }
  __syncwarp();
  // bitonic sort of the PN sample keys by one warp: in registers (KPL consecutive keys per lane, shuffles for the
  // cross-lane stages) when PN is 32..256, in shared memory otherwise
  if (PN == 32) warp_bitonic_sort_regs<1>(s_smp, lane);
  else if (PN == 64) warp_bitonic_sort_regs<2>(s_smp, lane);
  else if (PN == 128) warp_bitonic_sort_regs<4>(s_smp, lane);
  else if (PN == 256) warp_bitonic_sort_regs<8>(s_smp, lane);
  else {
    for (int k = 2; k <= PN; k <<= 1) {
      for (int j = k >> 1, lj = 31 - __clz(k >> 1); j > 0; j >>= 1, --lj) {
        for (int t = lane; t < PN / 2; t += 32) {
          const int i = ((t >> lj) << (lj + 1)) | (t & (j - 1));   // t-th compare-exchange pair of this stage
          const int l = i + j;
          const bool up = ((i & k) == 0);
          const float a = s_smp[i], b = s_smp[l];
          if ((a > b) == up) { s_smp[i] = b; s_smp[l] = a; }
        }
        __syncwarp();
      }
    }
  }
  // merge by rank.  Precondition: z ascending and NaN-free (the stratified coarse samples are).  The output row is
  // pre-filled with NaN so that a violated precondition (ranks collide, slots stay unwritten) is visible downstream
  // -- the compositing kernel's NaN flag fires -- instead of leaking stale shared memory into z_fine.
  for (int j = lane; j < S + N; j += 32) s_out[j] = __int_as_float(0x7fc00000);
  __syncwarp();
  for (int i = lane; i < S; i += 32) {
    const float v = s_z[i];
    int lo = 0, hi = N;                       // #{samples < v}
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_smp[mid] < v) lo = mid + 1; else hi = mid; }
    s_out[i + lo] = v;
  }
  for (int k = lane; k < N; k += 32) {
    const float v = s_smp[k];
    int lo = 0, hi = S;                       // #{z <= v}
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_z[mid] <= v) lo = mid + 1; else hi = mid; }
    s_out[k + lo] = v;
  }
  __syncwarp();
  for (int j = lane; j < S + N; j += 32) z_fine[r * (S + N) + j] = s_out[j];
}

}  // namespace lush

using namespace lush;

extern "C" int lush_sample_pdf(const float* bins, const float* weights, const float* cdf_in,
                               const float* u, int64_t u_stride, float* samples, int64_t* inds_out,
                               int64_t R, int nb, int N, void* stream) {
  LUSH_REQUIRE(bins && u && samples && (weights || cdf_in), "sample_pdf: null pointer");
  LUSH_REQUIRE(R >= 0 && nb >= 2 && N >= 1, "sample_pdf: need nb >= 2 bins and N >= 1 samples");
  LUSH_REQUIRE(nb <= 4096, "sample_pdf: nb > 4096 bins is not supported");
  if (R == 0) return 0;
  int per_warp = 3 * nb;
  size_t smem = (size_t)kSamplerWarps * per_warp * sizeof(float);
  if (smem > 48 * 1024) {
    int rc = check_cuda(cudaFuncSetAttribute(sample_pdf_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "sample_pdf: smem attr");
    if (rc) return rc;
  }
  sample_pdf_kernel<<<(unsigned)ceil_div(R, kSamplerWarps), kSamplerWarps * 32, smem,
                      (cudaStream_t)stream>>>(bins, weights, cdf_in, u, u_stride, samples, inds_out,
                                              R, nb, N, per_warp);
  LUSH_LAUNCH_CHECK("sample_pdf");
  return 0;
}

extern "C" int lush_sample_merge(const float* z, const float* weights, const float* cdf_in, const float* u,
                                 int64_t u_stride, float* z_fine, float* z_samples, float* z_std,
                                 int64_t R, int S, int N, void* stream) {
  LUSH_REQUIRE(z && (weights || cdf_in) && u && z_fine && z_samples, "sample_merge: null pointer");
  LUSH_REQUIRE(R >= 0 && S >= 3 && N >= 1, "sample_merge: need S >= 3 and N >= 1");
  LUSH_REQUIRE(S + N <= 2048, "sample_merge: S+N > 2048 is not supported");
  if (R == 0) return 0;
  int P = 1;                      // sort width of the N new samples
  while (P < N) P <<= 1;
  int per_warp = 3 * (S - 1) + S + P + (S + N);
  size_t smem = (size_t)kSamplerWarps * per_warp * sizeof(float);
  if (smem > 48 * 1024) {
    int rc = check_cuda(cudaFuncSetAttribute(sample_merge_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "sample_merge: smem attr");
    if (rc) return rc;
  }
  sample_merge_kernel<<<(unsigned)ceil_div(R, kSamplerWarps), kSamplerWarps * 32, smem,
                        (cudaStream_t)stream>>>(z, weights, cdf_in, u, u_stride, z_fine, z_samples,
                                                z_std, R, S, N, P, per_warp);
  LUSH_LAUNCH_CHECK("sample_merge");
  return 0;
}
