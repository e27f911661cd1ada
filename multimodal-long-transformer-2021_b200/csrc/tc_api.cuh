// Interface between the ABI layer and the tcgen05 (sm_100a tensor-core) kernels.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

#include "../../include/mlt_attn.h"
#include "mlt_common.cuh"

namespace mlt {

// True when the tcgen05 forward kernel can run this problem (bf16, d == 64, R <= 64, 16-byte
// aligned strides, TMA encode entry point available).
bool tc_fwd_args_supported(const FwdArgs& a, int dtype, int d);
// Enqueues the forward kernel; returns 0, an MLT_ERR_* (< 0) or a cudaError_t (> 0).
int tc_launch_fwd(const FwdArgs& a, cudaStream_t st);

}  // namespace mlt
