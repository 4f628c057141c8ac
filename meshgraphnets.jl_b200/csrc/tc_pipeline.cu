#include "tc.cuh"
namespace mgn {
int32_t tc_workspace_bytes(const mgn_model*, const mgn_graph*, bool, size_t*) {
  return fail(MGN_ERR_UNSUPPORTED, "MGN_COMPUTE_BF16 not built yet");
}
int32_t tc_forward(const mgn_model*, const mgn_graph*, const float*, const float*, const float*,
                   float*, void*, size_t, bool, cudaStream_t) {
  return fail(MGN_ERR_UNSUPPORTED, "MGN_COMPUTE_BF16 not built yet");
}
int32_t tc_backward(const mgn_model*, const mgn_graph*, const float*, const float*, const float*,
                    const float*, float*, float*, void*, size_t, cudaStream_t) {
  return fail(MGN_ERR_UNSUPPORTED, "MGN_COMPUTE_BF16 not built yet");
}
}  // namespace mgn
