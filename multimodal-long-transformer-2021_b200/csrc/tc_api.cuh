// Interface between the ABI layer and the tcgen05 (sm_100a tensor-core) kernels.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

#include "../../include/mlt_attn.h"
#include "mlt_common.cuh"

#include <mutex>

namespace mlt {

// Runs `fn` (returning 0 on success) once per CUDA device, thread-safe: kernel function attributes
// are per device, and the library is entered from arbitrary threads (TF's inter-op pool).  A failed
// attempt is retried by the next caller.
class PerDeviceOnce {
 public:
  template <typename F>
  int run(F&& fn) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMax) return (int)cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lk(mu_);
    if (done_[dev]) return 0;
    const int e = fn();
    if (e == 0) done_[dev] = true;
    return e;
  }

 private:
  static constexpr int kMax = 64;
  std::mutex mu_;
  bool done_[kMax] = {};
};

// [B, len, H, 64] bf16 view addressable by the TMA path (16-byte base / strides, broadcast strides
// only over extents of 1).
bool tc_t4_ok(const T4& t, int B, int len, int H);

// True when the tcgen05 forward kernel can run this problem (bf16, d == 64, R <= 64, 16-byte
// aligned strides, TMA encode entry point available).
bool tc_fwd_args_supported(const FwdArgs& a, int dtype, int d);
// Enqueues the forward kernel; returns 0, an MLT_ERR_* (< 0) or a cudaError_t (> 0).
int tc_launch_fwd(const FwdArgs& a, cudaStream_t st);


// ---- gl2: persistent kernels specialised for compact global-local attention (gl2_*.cu) ----------
// Long rows, forward: bf16, d = 64, example-id masks, 1-D band ids + sentence cross ids, R <= 32, no dropout.
bool gl2_fwd_long_supported(const FwdArgs& a, int dtype, int d);
int gl2_launch_fwd_long(const FwdArgs& a, cudaStream_t st);
// Long rows, query-centric backward (dQ, table-gradient partials, row records for the key-centric pass).
// `rowstat` / `rec_ws` are the workspace areas tc_launch_bwd_q carves (tc_bwd_prep_kernel has filled rowstat).
bool gl2_bwd_q_long_supported(const BwdQArgs& a, int dtype, int d);
int gl2_launch_bwd_q_long(const BwdQArgs& a, const float4* rowstat, float* rec_ws, int lp, int rw, cudaStream_t st);

// ---- backward (tc_bwd.cu) -------------------------------------------------------------------
// Extra workspace (bytes) of one row set: rowstat [B,H,Lpad] float4 + allrel [B,H,Lpad,R4] f32.
size_t tc_bwd_rows_ws_bytes(int B, int H, int len, int R);
bool tc_bwd_q_supported(const BwdQArgs& a, int dtype, int d);
// Query-centric pass: dq, dallrel (a.dallrel, [B,H,Lq,R]); publishes rowstat / allrel into `ws`.
int tc_launch_bwd_q(const BwdQArgs& a, void* ws, cudaStream_t st, bool allow_gl2 = false);
// Key-centric pass: dk, dv.  ws[s] is the workspace tc_launch_bwd_q filled for query source s.
int tc_launch_bwd_kv(const BwdKVArgs& a, void* const ws[2], cudaStream_t st);

}  // namespace mlt
