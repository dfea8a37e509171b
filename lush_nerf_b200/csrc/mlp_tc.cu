// placeholder until the tcgen05 path lands (replaced below in history)
#include "mlp_common.cuh"
using namespace lush;
bool lush_mlp_tc_supported(const MlpLayout& L) { (void)L; return false; }
size_t lush_mlp_tc_stash_bytes(const MlpLayout&, int64_t, int) { return 0; }
size_t lush_mlp_tc_workspace_bytes(const MlpLayout&, int64_t, int, int) { return 0; }
int lush_mlp_tc_fwd(const MlpLayout&, const lush_mlp_desc*, const float*, const float*, int, const float*,
                    float*, void*, void*, size_t, int64_t, int, cudaStream_t) { return LUSH_EUNSUPPORTED; }
int lush_mlp_tc_bwd(const MlpLayout&, const lush_mlp_desc*, const float*, const float*, int, const float*,
                    const float*, const void*, float*, float*, void*, size_t, int64_t, int, cudaStream_t) {
  return LUSH_EUNSUPPORTED;
}
