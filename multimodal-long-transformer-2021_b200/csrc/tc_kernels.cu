// tcgen05 / TMEM kernels (bring-up in progress): until they are parity-green the ABI reports
// "not supported" here and every request runs on the CUDA-core kernels.
#include "tc_api.cuh"

namespace mlt {

bool tc_gl_supported(const mlt_gl_params*) { return false; }
bool tc_dense_supported(const mlt_dense_params*) { return false; }
size_t tc_gl_workspace_bytes(const mlt_gl_params*, int) { return 0; }
size_t tc_dense_workspace_bytes(const mlt_dense_params*, int) { return 0; }
int tc_gl_fwd(const mlt_gl_params*, cudaStream_t) { return MLT_ERR_UNSUPPORTED; }
int tc_gl_bwd(const mlt_gl_params*, const mlt_gl_grads*, cudaStream_t) { return MLT_ERR_UNSUPPORTED; }
int tc_dense_fwd(const mlt_dense_params*, cudaStream_t) { return MLT_ERR_UNSUPPORTED; }
int tc_dense_bwd(const mlt_dense_params*, const mlt_dense_grads*, cudaStream_t) { return MLT_ERR_UNSUPPORTED; }

}  // namespace mlt
