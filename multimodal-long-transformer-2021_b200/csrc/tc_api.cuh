// Interface between the ABI layer and the tcgen05 (sm_100a tensor-core) kernels.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

#include "../../include/mlt_attn.h"

namespace mlt {

// True when the tcgen05 path can run these parameters (bf16, d == 64, R <= 64, ...).
bool tc_gl_supported(const mlt_gl_params* p);
bool tc_dense_supported(const mlt_dense_params* p);
size_t tc_gl_workspace_bytes(const mlt_gl_params* p, int bwd);
size_t tc_dense_workspace_bytes(const mlt_dense_params* p, int bwd);
int tc_gl_fwd(const mlt_gl_params* p, cudaStream_t st);
int tc_gl_bwd(const mlt_gl_params* p, const mlt_gl_grads* g, cudaStream_t st);
int tc_dense_fwd(const mlt_dense_params* p, cudaStream_t st);
int tc_dense_bwd(const mlt_dense_params* p, const mlt_dense_grads* g, cudaStream_t st);

}  // namespace mlt
