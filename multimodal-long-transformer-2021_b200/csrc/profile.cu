#include "profile.cuh"

#include <atomic>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/mlt_attn.h"

namespace mlt {
namespace {
std::atomic<int> g_enabled{0};
std::atomic<long long> g_launches{0};
std::mutex g_mu;
struct Rec {
  char name[48];
  double flops, bytes;
  cudaEvent_t beg, end;
};
std::vector<Rec> g_recs;
}  // namespace

bool profile_enabled() { return g_enabled.load(std::memory_order_relaxed) != 0; }
void profile_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

ProfileScope::ProfileScope(const char* name, double flops, double bytes, cudaStream_t st, int launches)
    : slot_(-1), st_(st) {
  profile_count_launch(launches);
  if (!profile_enabled()) return;
  Rec r{};
  std::strncpy(r.name, name, sizeof(r.name) - 1);
  r.flops = flops;
  r.bytes = bytes;
  if (cudaEventCreate(&r.beg) != cudaSuccess || cudaEventCreate(&r.end) != cudaSuccess) return;
  cudaEventRecord(r.beg, st);
  std::lock_guard<std::mutex> lk(g_mu);
  g_recs.push_back(r);
  slot_ = (int)g_recs.size() - 1;
}

ProfileScope::~ProfileScope() {
  if (slot_ < 0) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (slot_ < (int)g_recs.size()) cudaEventRecord(g_recs[slot_].end, st_);
}

}  // namespace mlt

extern "C" {

int mlt_profile_enable(int on) {
  mlt::g_enabled.store(on ? 1 : 0);
  return MLT_OK;
}

long long mlt_launch_count(void) { return mlt::g_launches.load(); }

int mlt_profile_read(mlt_kernel_time* out, int max_entries) {
  std::lock_guard<std::mutex> lk(mlt::g_mu);
  int n = 0;
  for (auto& r : mlt::g_recs) {
    float ms = 0.f;
    cudaEventSynchronize(r.end);
    cudaEventElapsedTime(&ms, r.beg, r.end);
    if (out && n < max_entries) {
      std::memcpy(out[n].name, r.name, sizeof(out[n].name));
      out[n].ms = ms;
      out[n].flops = r.flops;
      out[n].bytes = r.bytes;
      ++n;
    }
    cudaEventDestroy(r.beg);
    cudaEventDestroy(r.end);
  }
  mlt::g_recs.clear();
  return n;
}

}  // extern "C"
