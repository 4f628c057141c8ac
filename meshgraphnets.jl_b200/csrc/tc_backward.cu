// Tensor-core (MGN_COMPUTE_BF16) backward pass.
#include "tc.cuh"

namespace mgn {

int32_t tc_backward_scratch_bytes(const mgn_model*, const mgn_graph*, size_t* bytes) {
  *bytes = 0;
  return MGN_OK;
}

int32_t tc_backward(const mgn_model*, const mgn_graph*, const float*, const float*, const float*, const float*,
                    float*, float*, void*, size_t, cudaStream_t) {
  return fail(MGN_ERR_UNSUPPORTED, "mgn_backward: MGN_COMPUTE_BF16 backward is not built yet");
}

}  // namespace mgn
