// tcgen05 (5th-gen tensor core) path: bf16 operands, fp32 accumulation in TMEM.
#pragma once
#include "common.cuh"

namespace mgn {
int32_t tc_workspace_bytes(const mgn_model* m, const mgn_graph* g, bool training, size_t* bytes);
int32_t tc_forward(const mgn_model* m, const mgn_graph* g, const float* params, const float* nf,
                   const float* ef, float* out, void* ws, size_t ws_bytes, bool training,
                   cudaStream_t st);
int32_t tc_backward(const mgn_model* m, const mgn_graph* g, const float* params, const float* nf,
                    const float* ef, const float* dout, float* dparams, float* dnf, void* ws,
                    size_t ws_bytes, cudaStream_t st);
}  // namespace mgn
