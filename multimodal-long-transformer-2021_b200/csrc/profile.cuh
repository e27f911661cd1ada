// Optional per-kernel timing (bench / profiling facility, off by default).
// When enabled through mlt_profile_enable(1) every kernel launch of the library is bracketed
// by CUDA events recorded on the CALLER's stream; mlt_profile_read() synchronises those events
// and returns (name, ms, algorithmic flops, algorithmic bytes) per launch.  When disabled the
// only cost is one relaxed atomic load per launch.
#pragma once

#include <cuda_runtime.h>

namespace mlt {

bool profile_enabled();
void profile_count_launch(int n = 1);

class ProfileScope {
 public:
  ProfileScope(const char* name, double flops, double bytes, cudaStream_t st, int launches = 1);
  ~ProfileScope();
  ProfileScope(const ProfileScope&) = delete;
  ProfileScope& operator=(const ProfileScope&) = delete;

 private:
  int slot_;
  cudaStream_t st_;
};

}  // namespace mlt
