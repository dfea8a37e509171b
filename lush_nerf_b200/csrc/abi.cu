// C-ABI plumbing: error string, launch counter, MLP entry points dispatching on `impl`.
#include <atomic>
#include <stdarg.h>
#include <string.h>
#include "mlp_common.cuh"

namespace lush {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace lush

using namespace lush;

// implemented in mlp_simt.cu
size_t lush_mlp_simt_stash_bytes(const MlpLayout& L, int64_t R, int S);
size_t lush_mlp_simt_workspace_bytes(const MlpLayout& L, int64_t R, int S, int training);
int lush_mlp_simt_fwd(const MlpLayout& L, const lush_mlp_desc* desc, const float* params,
                      const float* ray_batch, int ray_stride, const float* z, float* out, void* stash,
                      void* workspace, size_t workspace_bytes, int64_t R, int S, cudaStream_t st);
int lush_mlp_simt_bwd(const MlpLayout& L, const lush_mlp_desc* desc, const float* params,
                      const float* ray_batch, int ray_stride, const float* z, const float* g_out,
                      const void* stash, float* g_params, float* g_ray_batch, void* workspace,
                      size_t workspace_bytes, int64_t R, int S, cudaStream_t st);
// implemented in mlp_tc.cu
bool lush_mlp_tc_supported(const MlpLayout& L);
bool lush_mlp_tc_training_supported();
size_t lush_mlp_tc_stash_bytes(const MlpLayout& L, int64_t R, int S);
size_t lush_mlp_tc_workspace_bytes(const MlpLayout& L, int64_t R, int S, int training);
int lush_mlp_tc_fwd(const MlpLayout& L, const lush_mlp_desc* desc, const float* params,
                    const float* ray_batch, int ray_stride, const float* z, float* out, void* stash,
                    void* workspace, size_t workspace_bytes, int64_t R, int S, cudaStream_t st);
int lush_mlp_tc_bwd(const MlpLayout& L, const lush_mlp_desc* desc, const float* params,
                    const float* ray_batch, int ray_stride, const float* z, const float* g_out,
                    const void* stash, float* g_params, float* g_ray_batch, void* workspace,
                    size_t workspace_bytes, int64_t R, int S, cudaStream_t st);

static bool layout_ok(const lush_mlp_desc* desc, MlpLayout& L, const char* who) {
  if (!make_layout(desc, L)) {
    set_error("%s: invalid lush_mlp_desc", who);
    return false;
  }
  if (L.skip == L.D - 1) {
    // the reference itself fails here (concat after the last layer feeds a W-wide head, helpers:400-407)
    set_error("%s: skip == D-1 is not a valid NeRF topology", who);
    return false;
  }
  if (L.skip >= L.D) L.skip = -1;
  if (desc->out_mode != 0 && desc->out_mode != 1) {
    set_error("%s: out_mode must be 0 or 1", who);
    return false;
  }
  return true;
}

extern "C" {

int lush_abi_version(void) { return LUSH_ABI_VERSION; }
const char* lush_last_error(void) { return g_err; }
int64_t lush_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int64_t lush_mlp_param_count(const lush_mlp_desc* desc) {
  MlpLayout L;
  if (!layout_ok(desc, L, "mlp_param_count")) return LUSH_EINVAL;
  return L.total;
}

int lush_mlp_impl_supported(const lush_mlp_desc* desc, int impl) {
  MlpLayout L;
  if (!layout_ok(desc, L, "mlp_impl_supported")) return 0;
  if (impl == 0) return 3;
  if (impl == 1) return lush_mlp_tc_supported(L) ? (lush_mlp_tc_training_supported() ? 3 : 1) : 0;
  return 0;
}

size_t lush_mlp_workspace_bytes(const lush_mlp_desc* desc, int64_t R, int S, int impl, int training) {
  MlpLayout L;
  if (!layout_ok(desc, L, "mlp_workspace_bytes") || R < 0 || S < 1) return 0;
  if (impl == 1 && lush_mlp_tc_supported(L)) return lush_mlp_tc_workspace_bytes(L, R, S, training);
  return lush_mlp_simt_workspace_bytes(L, R, S, training);
}

size_t lush_mlp_stash_bytes(const lush_mlp_desc* desc, int64_t R, int S, int impl) {
  MlpLayout L;
  if (!layout_ok(desc, L, "mlp_stash_bytes") || R < 0 || S < 1) return 0;
  if (impl == 1 && lush_mlp_tc_supported(L)) return lush_mlp_tc_stash_bytes(L, R, S);
  return lush_mlp_simt_stash_bytes(L, R, S);
}

int lush_mlp_fwd(const lush_mlp_desc* desc, const float* params, const float* ray_batch,
                 int ray_stride, const float* z, float* out, void* stash, void* workspace,
                 size_t workspace_bytes, int64_t R, int S, int impl, void* stream) {
  MlpLayout L;
  if (!layout_ok(desc, L, "mlp_fwd")) return LUSH_EINVAL;
  LUSH_REQUIRE(params && ray_batch && z && out, "mlp_fwd: null pointer");
  LUSH_REQUIRE(R >= 0 && S >= 1 && ray_stride >= 11, "mlp_fwd: bad R/S/ray_stride");
  LUSH_REQUIRE(impl == 0 || impl == 1, "mlp_fwd: impl must be 0 (fp32 SIMT) or 1 (tcgen05 fp16)");
  if (R == 0) return 0;
  if (impl == 1) {
    if (!lush_mlp_tc_supported(L)) {
      set_error("mlp_fwd: tcgen05 path supports W=256, skip<D-1 scene MLPs only (got W=%d D=%d)", L.W, L.D);
      return LUSH_EUNSUPPORTED;
    }
    return lush_mlp_tc_fwd(L, desc, params, ray_batch, ray_stride, z, out, stash, workspace,
                           workspace_bytes, R, S, (cudaStream_t)stream);
  }
  return lush_mlp_simt_fwd(L, desc, params, ray_batch, ray_stride, z, out, stash, workspace,
                           workspace_bytes, R, S, (cudaStream_t)stream);
}

int lush_mlp_bwd(const lush_mlp_desc* desc, const float* params, const float* ray_batch,
                 int ray_stride, const float* z, const float* g_out, const void* stash,
                 float* g_params, float* g_ray_batch, void* workspace, size_t workspace_bytes,
                 int64_t R, int S, int impl, void* stream) {
  MlpLayout L;
  if (!layout_ok(desc, L, "mlp_bwd")) return LUSH_EINVAL;
  LUSH_REQUIRE(params && ray_batch && z && g_out && stash && g_params, "mlp_bwd: null pointer");
  LUSH_REQUIRE(R >= 0 && S >= 1 && ray_stride >= 11, "mlp_bwd: bad R/S/ray_stride");
  LUSH_REQUIRE(impl == 0 || impl == 1, "mlp_bwd: impl must be 0 (fp32 SIMT) or 1 (tcgen05 fp16)");
  if (R == 0) return 0;
  if (impl == 1) {
    if (!lush_mlp_tc_supported(L)) {
      set_error("mlp_bwd: tcgen05 path supports W=256, skip<D-1 scene MLPs only (got W=%d D=%d)", L.W, L.D);
      return LUSH_EUNSUPPORTED;
    }
    return lush_mlp_tc_bwd(L, desc, params, ray_batch, ray_stride, z, g_out, stash, g_params,
                           g_ray_batch, workspace, workspace_bytes, R, S, (cudaStream_t)stream);
  }
  return lush_mlp_simt_bwd(L, desc, params, ray_batch, ray_stride, z, g_out, stash, g_params,
                           g_ray_batch, workspace, workspace_bytes, R, S, (cudaStream_t)stream);
}

}  // extern "C"
