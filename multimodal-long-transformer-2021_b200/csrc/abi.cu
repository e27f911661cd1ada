// extern "C" entry points of include/mlt_attn.h: argument validation, workspace carving and
// dispatch to the SIMT (simt_kernels.cu) or tcgen05 (tc_*.cu) kernels.  No allocation, no
// retained pointers, no exceptions across the boundary.

#include "../../include/mlt_attn.h"

#include "mlt_common.cuh"
#include "profile.cuh"
#include "tc_api.cuh"

#include <cstdio>
#include <mutex>

namespace {

using namespace mlt;

constexpr size_t kAlign = 256;
inline size_t align_up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

inline T4 to_t4(const mlt_tensor4& t) { return T4{t.ptr, t.stride_b, t.stride_l, t.stride_h}; }

inline int elem_size(int dtype) { return dtype == MLT_F32 ? 4 : 2; }

// ---- work accounting (algorithmic, DESIGN.md "Work accounting") --------------------------
inline double band_pairs(double L, double r) {
  if (L > r) return L * (2 * r + 1) - r * (r + 1);
  return L * L;
}
// forward of one row set: QK^T + PV over `pairs` attended pairs + q.E^T over R ids
inline double fwd_flops(double bh, double pairs, double d, double R, double lq) {
  return bh * (4 * d * pairs + 2 * d * R * lq);
}
inline double qkv_bytes(double bh, double rows, double d, int dtype, double ntensors) {
  return bh * rows * d * elem_size(dtype) * ntensors;
}

int check_t4(const mlt_tensor4& t, int dtype, int d) {
  if (t.ptr == nullptr) return MLT_ERR_NULL;
  // kernels move 4 elements (SIMT) / 16 bytes (tcgen05) per access
  if (t.stride_b % 4 || t.stride_l % 4 || t.stride_h % 4) return MLT_ERR_STRIDE;
  if (reinterpret_cast<uintptr_t>(t.ptr) % 16) return MLT_ERR_STRIDE;
  if (t.stride_h < d && t.stride_h != 0) return MLT_ERR_STRIDE;
  (void)dtype;
  return MLT_OK;
}

#define MLT_TRY(expr)            \
  do {                           \
    int _e = (expr);             \
    if (_e != MLT_OK) return _e; \
  } while (0)
#define MLT_CUDA(expr)                  \
  do {                                  \
    cudaError_t _e = (expr);            \
    if (_e != cudaSuccess) return (int)_e; \
  } while (0)

struct RowWs {  // backward workspace of one row set
  float* delta;
  float* allrel;
  float* dallrel;
  float* partial;
  float* partial_bias;
};

size_t row_ws_bytes(int B, int H, int len, int R, int d) {
  const int nchunk = simt_table_grad_chunks(len);
  size_t n = 0;
  n += align_up(sizeof(float) * (size_t)B * H * len);
  n += 2 * align_up(sizeof(float) * (size_t)B * H * len * (R > 0 ? R : 1));
  n += align_up(sizeof(float) * (size_t)B * nchunk * H * (R > 0 ? R : 1) * d);
  n += align_up(sizeof(float) * (size_t)B * nchunk * H * (R > 0 ? R : 1));
  return n;
}

RowWs carve_row_ws(char*& p, int B, int H, int len, int R, int d) {
  const int nchunk = simt_table_grad_chunks(len);
  const int r1 = R > 0 ? R : 1;
  RowWs w;
  w.delta = reinterpret_cast<float*>(p);
  p += align_up(sizeof(float) * (size_t)B * H * len);
  w.allrel = reinterpret_cast<float*>(p);
  p += align_up(sizeof(float) * (size_t)B * H * len * r1);
  w.dallrel = reinterpret_cast<float*>(p);
  p += align_up(sizeof(float) * (size_t)B * H * len * r1);
  w.partial = reinterpret_cast<float*>(p);
  p += align_up(sizeof(float) * (size_t)B * nchunk * H * r1 * d);
  w.partial_bias = reinterpret_cast<float*>(p);
  p += align_up(sizeof(float) * (size_t)B * nchunk * H * r1);
  return w;
}

// dropout_p in [0, 1); thr = floor(p * 2^32) (0 = off)
int check_dropout(float p) { return (p >= 0.f && p < 1.f) ? MLT_OK : MLT_ERR_DROPOUT; }
Dropout make_dropout(float p, uint64_t seed, uint32_t rowset) {
  Dropout d{};
  double t = (double)p * 4294967296.0;
  if (t > 4294967295.0) t = 4294967295.0;
  d.thr = p > 0.f ? (uint32_t)t : 0u;
  d.inv_keep = 1.f / (1.f - p);
  d.seed_lo = (uint32_t)seed;
  d.seed_hi = (uint32_t)(seed >> 32);
  d.rowset = rowset;
  return d;
}

Side explicit_side(const int32_t* mask, const int32_t* ids, int lq, int width) {
  Side s{};
  s.mask_rule = mask ? MR_EXPLICIT : MR_NONE;
  s.id_rule = ids ? IDR_EXPLICIT : IDR_NONE;
  s.mask = mask;
  s.ids = ids;
  s.sb = (int64_t)lq * width;
  s.sq = width;
  return s;
}

// ---- contract (A) ------------------------------------------------------------------------

int validate_dense(const mlt_dense_params* p) {
  if (!p) return MLT_ERR_NULL;
  if (p->abi_version != MLT_ABI_VERSION) return MLT_ERR_UNSUPPORTED;
  if (p->dtype != MLT_F32 && p->dtype != MLT_BF16) return MLT_ERR_DTYPE;
  if (p->B <= 0 || p->Lq <= 0 || p->Lk <= 0 || p->H <= 0 || p->d <= 0 || p->R < 0) return MLT_ERR_SHAPE;
  if (p->R > 64) return MLT_ERR_UNSUPPORTED;
  if (!simt_supports_head_dim(p->d)) return MLT_ERR_UNSUPPORTED;
  MLT_TRY(check_dropout(p->dropout_p));
  MLT_TRY(check_t4(p->q, p->dtype, p->d));
  MLT_TRY(check_t4(p->k, p->dtype, p->d));
  MLT_TRY(check_t4(p->v, p->dtype, p->d));
  MLT_TRY(check_t4(p->out, p->dtype, p->d));
  if (!p->stats) return MLT_ERR_NULL;
  if ((p->tables.emb == nullptr) != (p->tables.bias == nullptr)) return MLT_ERR_NULL;
  if (p->R > 0 && !p->tables.emb) return MLT_ERR_NULL;
  if (p->side_mode == MLT_SIDE_COMPACT) {
    if (!p->q_example_ids || !p->k_example_ids) return MLT_ERR_NULL;
    if (p->id_layout.max_distance < 0 || p->id_layout.num_patch_per_row < 0) return MLT_ERR_SHAPE;
    if (p->id_layout.num_patch_per_row > 0 && p->id_layout.num_core_layers <= 0) return MLT_ERR_SHAPE;
  } else if (p->side_mode != MLT_SIDE_EXPLICIT) {
    return MLT_ERR_UNSUPPORTED;
  }
  return MLT_OK;
}

Side dense_side(const mlt_dense_params* p) {
  if (p->side_mode == MLT_SIDE_EXPLICIT) {
    Side s = explicit_side(p->att_mask, p->R > 0 ? p->relative_att_ids : nullptr, p->Lq, p->Lk);
    return s;
  }
  Side s{};
  s.mask_rule = MR_EXAMPLE_ID;
  s.q_eid = p->q_example_ids;
  s.k_eid = p->k_example_ids;
  s.q_len = p->Lq;
  s.k_len = p->Lk;
  s.max_distance = p->id_layout.max_distance;
  s.npr = p->id_layout.num_patch_per_row;
  s.core = p->id_layout.num_core_layers;
  s.id_rule = p->R == 0 ? IDR_NONE : (s.npr > 0 ? IDR_2D : IDR_1D);
  return s;
}

RowSet dense_rows(const mlt_dense_params* p) {
  return RowSet{to_t4(p->q), p->Lq, p->tables.emb, p->tables.bias, p->R};
}

// ---- contract (B) ------------------------------------------------------------------------

int validate_gl(const mlt_gl_params* p) {
  if (!p) return MLT_ERR_NULL;
  if (p->abi_version != MLT_ABI_VERSION) return MLT_ERR_UNSUPPORTED;
  if (p->dtype != MLT_F32 && p->dtype != MLT_BF16) return MLT_ERR_DTYPE;
  if (p->B <= 0 || p->L <= 0 || p->G <= 0 || p->H <= 0 || p->d <= 0 || p->R < 0 ||
      p->local_radius < 1)
    return MLT_ERR_SHAPE;
  if (p->R > 64) return MLT_ERR_UNSUPPORTED;
  if (!simt_supports_head_dim(p->d)) return MLT_ERR_UNSUPPORTED;
  MLT_TRY(check_dropout(p->dropout_p));
  const mlt_tensor4* ts[] = {&p->long_q, &p->long_k, &p->long_v, &p->global_q, &p->global_k,
                             &p->global_v, &p->long_out, &p->global_out};
  for (const mlt_tensor4* t : ts) MLT_TRY(check_t4(*t, p->dtype, p->d));
  if (!p->long_stats || !p->global_stats) return MLT_ERR_NULL;
  const mlt_rel_tables* tb[] = {&p->long_tables, &p->global_tables};
  for (const mlt_rel_tables* t : tb) {
    if ((t->emb == nullptr) != (t->bias == nullptr)) return MLT_ERR_NULL;
    if (p->R > 0 && !t->emb) return MLT_ERR_NULL;
  }
  if (p->side_mode == MLT_SIDE_COMPACT) {
    if (!p->long_example_ids || !p->global_example_ids) return MLT_ERR_NULL;
    if (p->R > 0 && !p->sentence_ids) return MLT_ERR_NULL;
    if (p->max_distance < 0) return MLT_ERR_SHAPE;
  } else if (p->side_mode != MLT_SIDE_EXPLICIT) {
    return MLT_ERR_UNSUPPORTED;
  }
  return MLT_OK;
}

enum GlBlock { L2L, L2G, G2G, G2L };

Side gl_side(const mlt_gl_params* p, GlBlock blk) {
  const int W = 2 * p->local_radius + 1;
  if (p->side_mode == MLT_SIDE_EXPLICIT) {
    const bool rel = p->R > 0;
    switch (blk) {
      case L2L: return explicit_side(p->l2l_att_mask, rel ? p->l2l_relative_att_ids : nullptr, p->L, W);
      case L2G: return explicit_side(p->l2g_att_mask, rel ? p->l2g_relative_att_ids : nullptr, p->L, p->G);
      case G2G: return explicit_side(p->g2g_att_mask, rel ? p->g2g_relative_att_ids : nullptr, p->G, p->G);
      default:  return explicit_side(p->g2l_att_mask, rel ? p->g2l_relative_att_ids : nullptr, p->G, p->L);
    }
  }
  Side s{};
  s.mask_rule = MR_EXAMPLE_ID;
  s.max_distance = p->max_distance;
  s.sent = p->sentence_ids;
  s.sent_len = p->L;
  const bool q_long = (blk == L2L || blk == L2G);
  const bool k_long = (blk == L2L || blk == G2L);
  s.q_eid = q_long ? p->long_example_ids : p->global_example_ids;
  s.k_eid = k_long ? p->long_example_ids : p->global_example_ids;
  s.q_len = q_long ? p->L : p->G;
  s.k_len = k_long ? p->L : p->G;
  if (p->R == 0) {
    s.id_rule = IDR_NONE;
  } else {
    s.id_rule = (blk == L2L || blk == G2G) ? IDR_1D : (blk == L2G ? IDR_CROSS_QSENT : IDR_CROSS_KSENT);
  }
  return s;
}

KeySeg make_seg(const mlt_tensor4& k, const mlt_tensor4& v, int len, int band, int radius,
                const Side& side, int col_base = 0) {
  KeySeg s{};
  s.k = to_t4(k);
  s.v = to_t4(v);
  s.len = len;
  s.band = band;
  s.radius = radius;
  s.side = side;
  s.col_base = col_base;
  return s;
}
QuerySource make_src(const RowSet& rows, const mlt_tensor4& d_out, const float* stats, const RowWs& ws, int band,
                     int radius, const Side& side, int col_base, const Dropout& drop) {
  QuerySource q{};
  q.rows = rows;
  q.d_out = to_t4(d_out);
  q.stats = stats;
  q.delta = ws.delta;
  q.allrel = ws.allrel;
  q.band = band;
  q.radius = radius;
  q.side = side;
  q.col_base = col_base;
  q.drop = drop;
  return q;
}

// ---- fork / join helper: runs independent kernels of one call on a second, library-owned
// stream so that the few long-running global-row tiles overlap the many long-row tiles.
// Plain event fork/join (CUDA-graph-capture friendly).  Disabled while per-kernel profiling is
// on (timings must stay attributable) and when event/stream creation failed.
struct ForkCtx {
  cudaStream_t s2 = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool ok = false, tried = false;
  std::mutex mu;   // held by the one call that is using this context (try_lock only: never blocks)
};
// A small pool of side-stream contexts per device.  A call takes any free context with try_lock; when
// all are busy (many host threads enqueueing on one device at once) it simply runs unforked on the
// caller's stream -- no call ever waits for another one.
constexpr int kMaxDevices = 64, kForkPool = 4;
ForkCtx g_fork[kMaxDevices][kForkPool];

ForkCtx* acquire_fork() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  for (int i = 0; i < kForkPool; ++i) {
    ForkCtx& f = g_fork[dev][i];
    if (!f.mu.try_lock()) continue;
    if (!f.tried) {
      f.tried = true;
      f.ok = cudaStreamCreateWithFlags(&f.s2, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&f.fork, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&f.join, cudaEventDisableTiming) == cudaSuccess;
    }
    if (f.ok) return &f;   // locked
    f.mu.unlock();
  }
  return nullptr;
}

// RAII: between construction and join() the side stream is ordered after everything enqueued on
// `st` so far; join() orders `st` after the side stream.  Owns one pool context for the enqueue window.
class ForkScope {
 public:
  explicit ForkScope(cudaStream_t st) : st_(st), f_(profile_enabled() ? nullptr : acquire_fork()) {
    if (f_) {
      if (cudaEventRecord(f_->fork, st_) != cudaSuccess || cudaStreamWaitEvent(f_->s2, f_->fork, 0) != cudaSuccess) {
        f_->mu.unlock();
        f_ = nullptr;
      }
    }
  }
  cudaStream_t side() const { return f_ ? f_->s2 : st_; }
  void join() {
    if (f_ && !joined_) {
      cudaEventRecord(f_->join, f_->s2);
      cudaStreamWaitEvent(st_, f_->join, 0);
      joined_ = true;
    }
  }
  ~ForkScope() {
    if (f_) {
      join();
      f_->mu.unlock();
    }
  }

 private:
  cudaStream_t st_;
  ForkCtx* f_;
  bool joined_ = false;
};

// Compact 2-D descriptors on the tcgen05 path: the ids depend on the two positions only, not on the
// batch element, so the library materialises ONE [Lq, Lk] int32 plane in the caller's workspace
// (side-input constructor kernel, ~1 MB at S = 512) and the kernels read it through the EXPL form
// with batch stride 0; the mask stays the example-id rule.  (The closed 2-D rule needs two integer
// divisions per element and otherwise runs through the generic per-element loop: 12x slower.)
size_t dense_ids_plane_bytes(const mlt_dense_params* p) {
  const bool plane = p->side_mode == MLT_SIDE_COMPACT && p->id_layout.num_patch_per_row > 0 && p->R > 0 &&
                     p->Lq == p->Lk;
  return plane ? align_up(sizeof(int32_t) * (size_t)p->Lq * p->Lk) : 0;
}
// The plane sits at the end of the workspace (mlt_dense_workspace_bytes); null when not applicable.
int32_t* dense_ids_plane(const mlt_dense_params* p, int bwd) {
  const size_t n = dense_ids_plane_bytes(p);
  const size_t total = mlt_dense_workspace_bytes(p, bwd);
  if (n == 0 || !p->workspace || p->workspace_bytes < total) return nullptr;
  char* base = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(p->workspace)));
  return reinterpret_cast<int32_t*>(base + (total - kAlign - n));
}
Side with_ids_plane(Side s, const int32_t* plane, int lk) {
  if (plane) {
    s.id_rule = IDR_EXPLICIT;
    s.ids = plane;
    s.sb = 0;      // every batch element reads the same plane
    s.sq = lk;
  }
  return s;
}

FwdArgs dense_fwd_args(const mlt_dense_params* p, const int32_t* ids_plane = nullptr) {
  FwdArgs a{};
  a.rows = dense_rows(p);
  a.seg[0] = make_seg(p->k, p->v, p->Lk, 0, 0, with_ids_plane(dense_side(p), ids_plane, p->Lk));
  a.nseg = 1;
  a.out = to_t4(p->out);
  a.stats = p->stats;
  a.B = p->B;
  a.H = p->H;
  a.scale = p->scale;
  a.neg = p->neg;
  a.drop = make_dropout(p->dropout_p, p->dropout_seed, 0);
  return a;
}

// long rows: band(l2l) (+) dense(l2g), long tables
FwdArgs gl_long_fwd_args(const mlt_gl_params* p) {
  FwdArgs a{};
  a.rows = RowSet{to_t4(p->long_q), p->L, p->long_tables.emb, p->long_tables.bias, p->R};
  a.seg[0] = make_seg(p->long_k, p->long_v, p->L, 1, p->local_radius, gl_side(p, L2L));
  a.seg[1] = make_seg(p->global_k, p->global_v, p->G, 0, 0, gl_side(p, L2G), p->L);
  a.nseg = 2;
  a.out = to_t4(p->long_out);
  a.stats = p->long_stats;
  a.B = p->B; a.H = p->H; a.scale = p->scale; a.neg = p->neg;
  a.drop = make_dropout(p->dropout_p, p->dropout_seed, 0);
  return a;
}

// global rows: dense(g2g) (+) dense(g2l), global tables
FwdArgs gl_global_fwd_args(const mlt_gl_params* p) {
  FwdArgs g{};
  g.rows = RowSet{to_t4(p->global_q), p->G, p->global_tables.emb, p->global_tables.bias, p->R};
  g.seg[0] = make_seg(p->global_k, p->global_v, p->G, 0, 0, gl_side(p, G2G));
  g.seg[1] = make_seg(p->long_k, p->long_v, p->L, 0, 0, gl_side(p, G2L), p->G);
  g.nseg = 2;
  g.out = to_t4(p->global_out);
  g.stats = p->global_stats;
  g.B = p->B; g.H = p->H; g.scale = p->scale; g.neg = p->neg;
  g.drop = make_dropout(p->dropout_p, p->dropout_seed, 1);
  return g;
}

bool gl_fwd_on_tc(const mlt_gl_params* p) {
  if (p->impl == MLT_IMPL_SIMT) return false;
  return tc_fwd_args_supported(gl_long_fwd_args(p), p->dtype, p->d) &&
         tc_fwd_args_supported(gl_global_fwd_args(p), p->dtype, p->d);
}
bool dense_fwd_on_tc(const mlt_dense_params* p) {
  if (p->impl == MLT_IMPL_SIMT) return false;
  return tc_fwd_args_supported(dense_fwd_args(p), p->dtype, p->d);
}

int launch_fwd(const FwdArgs& a, bool tc, int dtype, int d, const char* name, double flops,
               double bytes, cudaStream_t st, bool allow_gl2 = false) {
  char full[48];
  const bool gl2 = tc && allow_gl2 && gl2_fwd_long_supported(a, dtype, d);
  snprintf(full, sizeof(full), "%s_%s", gl2 ? "gl2" : (tc ? "tc" : "simt"), name);
  ProfileScope ps(full, flops, bytes, st);
  if (gl2) return gl2_launch_fwd_long(a, st);
  if (tc) return tc_launch_fwd(a, st);
  MLT_CUDA(simt_launch_fwd(a, dtype, d, st));
  return MLT_OK;
}

bool bwd_kv_on_tc(const BwdKVArgs& kv) {
  auto ok = [&](const T4& t, int len) { return tc_t4_ok(t, kv.B, len, kv.H); };
  if (!ok(kv.k, kv.len) || !ok(kv.v, kv.len) || !ok(kv.d_k, kv.len) || !ok(kv.d_v, kv.len)) return false;
  for (int s = 0; s < kv.nsrc; ++s)
    if (!ok(kv.src[s].rows.q, kv.src[s].rows.len) || !ok(kv.src[s].d_out, kv.src[s].rows.len)) return false;
  return true;
}

int launch_bwd_q(const BwdQArgs& a, bool tc, void* tc_ws, int dtype, int d, const char* name, double flops,
                 double bytes, cudaStream_t st, bool allow_gl2 = false) {
  char full[48];
  const bool gl2 = tc && allow_gl2 && gl2_bwd_q_long_supported(a, dtype, d);
  snprintf(full, sizeof(full), "%s_%s", gl2 ? "gl2" : (tc ? "tc" : "simt"), name);
  ProfileScope ps(full, flops, bytes, st, tc ? 2 : 1);   // tcgen05 path = row-record preprocess + main kernel
  if (tc) return tc_launch_bwd_q(a, tc_ws, st, gl2);
  MLT_CUDA(simt_launch_bwd_q(a, dtype, d, st));
  return MLT_OK;
}
int launch_bwd_kv(const BwdKVArgs& a, bool tc, void* const tc_ws[2], int dtype, int d, const char* name,
                  double flops, double bytes, cudaStream_t st) {
  char full[48];
  snprintf(full, sizeof(full), "%s_%s", tc ? "tc" : "simt", name);
  ProfileScope ps(full, flops, bytes, st);
  if (tc) return tc_launch_bwd_kv(a, tc_ws, st);
  MLT_CUDA(simt_launch_bwd_kv(a, dtype, d, st));
  return MLT_OK;
}

}  // namespace

// ============================================================================================
extern "C" {

int mlt_abi_version(void) { return MLT_ABI_VERSION; }

const char* mlt_strerror(int code) {
  switch (code) {
    case MLT_OK: return "ok";
    case MLT_ERR_NULL: return "mlt: required pointer is NULL";
    case MLT_ERR_SHAPE: return "mlt: non-positive or inconsistent dimension";
    case MLT_ERR_UNSUPPORTED: return "mlt: unsupported configuration for this build";
    case MLT_ERR_STRIDE: return "mlt: stride/alignment requirement violated (4 elements, 16 bytes)";
    case MLT_ERR_WORKSPACE: return "mlt: workspace missing or too small";
    case MLT_ERR_DTYPE: return "mlt: unknown dtype";
    case MLT_ERR_DROPOUT: return "mlt: dropout_p must lie in [0, 1)";
    default:
      if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
      return "mlt: unknown error";
  }
}

int mlt_dense_uses_tensor_cores(const mlt_dense_params* p) {
  if (validate_dense(p) != MLT_OK) return 0;
  return dense_fwd_on_tc(p) ? 1 : 0;
}
int mlt_gl_uses_tensor_cores(const mlt_gl_params* p) {
  if (validate_gl(p) != MLT_OK) return 0;
  return gl_fwd_on_tc(p) ? 1 : 0;
}

size_t mlt_dense_workspace_bytes(const mlt_dense_params* p, int bwd) {
  if (!p) return 0;
  size_t n = kAlign;
  if (bwd) n += row_ws_bytes(p->B, p->H, p->Lq, p->R, p->d) + tc_bwd_rows_ws_bytes(p->B, p->H, p->Lq, p->R);
  n += dense_ids_plane_bytes(p);
  return n;
}

size_t mlt_gl_workspace_bytes(const mlt_gl_params* p, int bwd) {
  if (!p) return 0;
  size_t n = kAlign;
  if (bwd)
    n += row_ws_bytes(p->B, p->H, p->L, p->R, p->d) + row_ws_bytes(p->B, p->H, p->G, p->R, p->d) +
         tc_bwd_rows_ws_bytes(p->B, p->H, p->L, p->R) + tc_bwd_rows_ws_bytes(p->B, p->H, p->G, p->R);
  return n;
}

int mlt_dense_rel_attn_fwd(const mlt_dense_params* p, void* cuda_stream) {
  MLT_TRY(validate_dense(p));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const bool tc = dense_fwd_on_tc(p);
  if ((p->impl == MLT_IMPL_TC || p->impl == MLT_IMPL_TC_GENERIC) && !tc) return MLT_ERR_UNSUPPORTED;
  const double bh = (double)p->B * p->H, pairs = (double)p->Lq * p->Lk;
  const int32_t* plane = nullptr;
  if (tc && dense_ids_plane_bytes(p)) {
    int32_t* w = dense_ids_plane(p, 0);
    if (!w) return MLT_ERR_WORKSPACE;
    MLT_TRY(mlt_build_dense_side_inputs(nullptr, 1, p->Lq, p->id_layout, nullptr, w, cuda_stream));
    plane = w;
  }
  return launch_fwd(dense_fwd_args(p, plane), tc, p->dtype, p->d, "fwd_dense",
                    fwd_flops(bh, pairs, p->d, p->R, p->Lq),
                    qkv_bytes(bh, 2.0 * p->Lq + 2.0 * p->Lk, p->d, p->dtype, 1), st);
}

int mlt_dense_rel_attn_bwd(const mlt_dense_params* p, const mlt_dense_grads* g, void* cuda_stream) {
  MLT_TRY(validate_dense(p));
  if (!g) return MLT_ERR_NULL;
  MLT_TRY(check_t4(g->d_out, p->dtype, p->d));
  MLT_TRY(check_t4(g->d_q, p->dtype, p->d));
  MLT_TRY(check_t4(g->d_k, p->dtype, p->d));
  MLT_TRY(check_t4(g->d_v, p->dtype, p->d));
  if (p->R > 0 && (!g->d_emb || !g->d_bias)) return MLT_ERR_NULL;
  if (!p->workspace || p->workspace_bytes < mlt_dense_workspace_bytes(p, 1)) return MLT_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  char* wp = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(p->workspace)));
  RowWs ws = carve_row_ws(wp, p->B, p->H, p->Lq, p->R, p->d);
  void* tcw[2] = {wp, wp};
  Side side = dense_side(p);

  BwdQArgs q{};
  q.rows = dense_rows(p);
  q.seg[0] = make_seg(p->k, p->v, p->Lk, 0, 0, side);
  q.nseg = 1;
  q.out = to_t4(p->out);
  q.d_out = to_t4(g->d_out);
  q.d_q = to_t4(g->d_q);
  q.stats = p->stats;
  q.delta = ws.delta;
  q.allrel = ws.allrel;
  q.dallrel = ws.dallrel;
  q.B = p->B; q.H = p->H; q.scale = p->scale; q.neg = p->neg;
  q.drop = make_dropout(p->dropout_p, p->dropout_seed, 0);
  BwdKVArgs kv{};
  kv.k = to_t4(p->k); kv.v = to_t4(p->v); kv.d_k = to_t4(g->d_k); kv.d_v = to_t4(g->d_v);
  kv.len = p->Lk;
  kv.src[0] = make_src(dense_rows(p), g->d_out, p->stats, ws, 0, 0, side, 0, q.drop);
  kv.nsrc = 1;
  kv.B = p->B; kv.H = p->H; kv.scale = p->scale; kv.neg = p->neg;

  const bool tc = p->impl != MLT_IMPL_SIMT && tc_bwd_q_supported(q, p->dtype, p->d) && bwd_kv_on_tc(kv);
  if ((p->impl == MLT_IMPL_TC || p->impl == MLT_IMPL_TC_GENERIC) && !tc) return MLT_ERR_UNSUPPORTED;
  const double bh = (double)p->B * p->H, pairs = (double)p->Lq * p->Lk;
  if (tc && dense_ids_plane_bytes(p)) {   // compact 2-D ids: one [Lq, Lk] plane, read through the EXPL form
    int32_t* w = dense_ids_plane(p, 1);
    if (!w) return MLT_ERR_WORKSPACE;
    MLT_TRY(mlt_build_dense_side_inputs(nullptr, 1, p->Lq, p->id_layout, nullptr, w, cuda_stream));
    side = with_ids_plane(side, w, p->Lk);
    q.seg[0].side = side;
    kv.src[0].side = side;
  }
  MLT_TRY(launch_bwd_q(q, tc, tcw[0], p->dtype, p->d, "bwd_q_dense",
                       bh * (4.0 * p->d * pairs + 2.0 * p->d * p->R * p->Lq),
                       qkv_bytes(bh, 4.0 * p->Lq + 2.0 * p->Lk, p->d, p->dtype, 1), st));
  MLT_TRY(launch_bwd_kv(kv, tc, tcw, p->dtype, p->d, "bwd_kv_dense", bh * 4.0 * p->d * pairs,
                        qkv_bytes(bh, 2.0 * p->Lq + 4.0 * p->Lk, p->d, p->dtype, 1), st));
  if (p->R > 0) {
    TableGradArgs t{to_t4(p->q), p->Lq, ws.dallrel, ws.partial, ws.partial_bias, g->d_emb, g->d_bias,
                    p->B, p->H, p->R, p->d, simt_table_grad_chunks(p->Lq), p->scale};
    ProfileScope ps("simt_table_grad_dense", bh * 2.0 * p->d * p->R * p->Lq,
                    qkv_bytes(bh, p->Lq, p->d, p->dtype, 1), st, 2);
    MLT_CUDA(simt_launch_table_grad(t, p->dtype, st));
  }
  return MLT_OK;
}

int mlt_gl_attn_fwd(const mlt_gl_params* p, void* cuda_stream) {
  MLT_TRY(validate_gl(p));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const bool tc = gl_fwd_on_tc(p);
  if ((p->impl == MLT_IMPL_TC || p->impl == MLT_IMPL_TC_GENERIC) && !tc) return MLT_ERR_UNSUPPORTED;
  const double bh = (double)p->B * p->H;
  const double pl = band_pairs(p->L, p->local_radius) + (double)p->L * p->G;
  const double pg = (double)p->G * (p->G + p->L);
  // global rows (few, long-running tiles) on the side stream, overlapping the long rows
  ForkScope fk(st);
  MLT_TRY(launch_fwd(gl_global_fwd_args(p), tc, p->dtype, p->d, "fwd_global_rows",
                     fwd_flops(bh, pg, p->d, p->R, p->G),
                     qkv_bytes(bh, 2.0 * p->L + 4.0 * p->G, p->d, p->dtype, 1), fk.side()));
  MLT_TRY(launch_fwd(gl_long_fwd_args(p), tc, p->dtype, p->d, "fwd_long_rows",
                     fwd_flops(bh, pl, p->d, p->R, p->L),
                     qkv_bytes(bh, 4.0 * p->L + 2.0 * p->G, p->d, p->dtype, 1), st,
                     p->impl != MLT_IMPL_TC_GENERIC));
  return MLT_OK;
}

int mlt_gl_attn_bwd(const mlt_gl_params* p, const mlt_gl_grads* g, void* cuda_stream) {
  MLT_TRY(validate_gl(p));
  if (!g) return MLT_ERR_NULL;
  const mlt_tensor4* ts[] = {&g->d_long_out, &g->d_global_out, &g->d_long_q, &g->d_long_k,
                             &g->d_long_v, &g->d_global_q, &g->d_global_k, &g->d_global_v};
  for (const mlt_tensor4* t : ts) MLT_TRY(check_t4(*t, p->dtype, p->d));
  if (p->R > 0 && (!g->d_long_emb || !g->d_long_bias || !g->d_global_emb || !g->d_global_bias))
    return MLT_ERR_NULL;
  if (!p->workspace || p->workspace_bytes < mlt_gl_workspace_bytes(p, 1)) return MLT_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  char* wp = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(p->workspace)));
  RowWs wl = carve_row_ws(wp, p->B, p->H, p->L, p->R, p->d);
  RowWs wg = carve_row_ws(wp, p->B, p->H, p->G, p->R, p->d);
  void* tc_wl = wp;
  wp += tc_bwd_rows_ws_bytes(p->B, p->H, p->L, p->R);
  void* tc_wg = wp;
  const RowSet long_rows{to_t4(p->long_q), p->L, p->long_tables.emb, p->long_tables.bias, p->R};
  const RowSet glob_rows{to_t4(p->global_q), p->G, p->global_tables.emb, p->global_tables.bias, p->R};
  const Side s_l2l = gl_side(p, L2L), s_l2g = gl_side(p, L2G), s_g2g = gl_side(p, G2G),
             s_g2l = gl_side(p, G2L);

  BwdQArgs ql{};
  ql.rows = long_rows;
  ql.seg[0] = make_seg(p->long_k, p->long_v, p->L, 1, p->local_radius, s_l2l);
  ql.seg[1] = make_seg(p->global_k, p->global_v, p->G, 0, 0, s_l2g, p->L);
  ql.nseg = 2;
  ql.drop = make_dropout(p->dropout_p, p->dropout_seed, 0);
  ql.out = to_t4(p->long_out); ql.d_out = to_t4(g->d_long_out); ql.d_q = to_t4(g->d_long_q);
  ql.stats = p->long_stats; ql.delta = wl.delta; ql.allrel = wl.allrel; ql.dallrel = wl.dallrel;
  ql.B = p->B; ql.H = p->H; ql.scale = p->scale; ql.neg = p->neg;
  ql.tg_partial = wl.partial; ql.tg_partial_bias = wl.partial_bias;
  BwdQArgs qg{};
  qg.rows = glob_rows;
  qg.seg[0] = make_seg(p->global_k, p->global_v, p->G, 0, 0, s_g2g);
  qg.seg[1] = make_seg(p->long_k, p->long_v, p->L, 0, 0, s_g2l, p->G);
  qg.nseg = 2;
  qg.drop = make_dropout(p->dropout_p, p->dropout_seed, 1);
  qg.out = to_t4(p->global_out); qg.d_out = to_t4(g->d_global_out); qg.d_q = to_t4(g->d_global_q);
  qg.stats = p->global_stats; qg.delta = wg.delta; qg.allrel = wg.allrel; qg.dallrel = wg.dallrel;
  qg.B = p->B; qg.H = p->H; qg.scale = p->scale; qg.neg = p->neg;
  qg.tg_partial = wg.partial; qg.tg_partial_bias = wg.partial_bias;
  // long keys: from long queries (band, l2l) and global queries (dense, g2l)
  BwdKVArgs kl{};
  kl.k = to_t4(p->long_k); kl.v = to_t4(p->long_v); kl.d_k = to_t4(g->d_long_k); kl.d_v = to_t4(g->d_long_v);
  kl.len = p->L;
  kl.src[0] = make_src(long_rows, g->d_long_out, p->long_stats, wl, 1, p->local_radius, s_l2l, 0, ql.drop);
  kl.src[1] = make_src(glob_rows, g->d_global_out, p->global_stats, wg, 0, 0, s_g2l, p->G, qg.drop);
  kl.nsrc = 2;
  kl.B = p->B; kl.H = p->H; kl.scale = p->scale; kl.neg = p->neg;
  // global keys: from long queries (dense, l2g) and global queries (dense, g2g)
  BwdKVArgs kg{};
  kg.k = to_t4(p->global_k); kg.v = to_t4(p->global_v); kg.d_k = to_t4(g->d_global_k); kg.d_v = to_t4(g->d_global_v);
  kg.len = p->G;
  kg.src[0] = make_src(long_rows, g->d_long_out, p->long_stats, wl, 0, 0, s_l2g, p->L, ql.drop);
  kg.src[1] = make_src(glob_rows, g->d_global_out, p->global_stats, wg, 0, 0, s_g2g, 0, qg.drop);
  kg.nsrc = 2;
  kg.B = p->B; kg.H = p->H; kg.scale = p->scale; kg.neg = p->neg;

  const bool tc = p->impl != MLT_IMPL_SIMT && tc_bwd_q_supported(ql, p->dtype, p->d) &&
                  tc_bwd_q_supported(qg, p->dtype, p->d) && bwd_kv_on_tc(kl) && bwd_kv_on_tc(kg);
  if ((p->impl == MLT_IMPL_TC || p->impl == MLT_IMPL_TC_GENERIC) && !tc) return MLT_ERR_UNSUPPORTED;
  void* ws_lg[2] = {tc_wl, tc_wg};

  const double bh = (double)p->B * p->H, dd = p->d, RR = p->R;
  const double p_l2l = band_pairs(p->L, p->local_radius), p_lg = (double)p->L * p->G,
               p_gg = (double)p->G * p->G;
  {
    ForkScope fk(st);
    MLT_TRY(launch_bwd_q(qg, tc, tc_wg, p->dtype, p->d, "bwd_q_global_rows",
                         bh * (4 * dd * (p_gg + p_lg) + 2 * dd * RR * p->G),
                         qkv_bytes(bh, 2.0 * p->L + 6.0 * p->G, p->d, p->dtype, 1), fk.side()));
  MLT_TRY(launch_bwd_q(ql, tc, tc_wl, p->dtype, p->d, "bwd_q_long_rows",
                       bh * (4 * dd * (p_l2l + p_lg) + 2 * dd * RR * p->L),
                       qkv_bytes(bh, 6.0 * p->L + 2.0 * p->G, p->d, p->dtype, 1), st,
                       p->impl != MLT_IMPL_TC_GENERIC));
  }
  // Main stream: the long-key kernel (longest).  Side stream: the global-key kernel, then the small
  // table-gradient kernels (they only need the bins of the query-centric pass).
  ForkScope fk2(st);
  cudaStream_t s2 = fk2.side();
  if (profile_enabled()) s2 = st;
  // the few long-running tiles of the global keys go first, the many short tiles of the long keys fill in
  MLT_TRY(launch_bwd_kv(kg, tc, ws_lg, p->dtype, p->d, "bwd_kv_global_keys", bh * 4 * dd * (p_lg + p_gg),
                        qkv_bytes(bh, 2.0 * p->L + 6.0 * p->G, p->d, p->dtype, 1), s2));
  MLT_TRY(launch_bwd_kv(kl, tc, ws_lg, p->dtype, p->d, "bwd_kv_long_keys", bh * 4 * dd * (p_l2l + p_lg),
                        qkv_bytes(bh, 6.0 * p->L + 2.0 * p->G, p->d, p->dtype, 1), st));
  if (p->R > 0) {
    TableGradArgs tl{to_t4(p->long_q), p->L, wl.dallrel, wl.partial, wl.partial_bias, g->d_long_emb,
                     g->d_long_bias, p->B, p->H, p->R, p->d, simt_table_grad_chunks(p->L), p->scale};
    {
      ProfileScope ps(tc ? "table_grad_reduce_long" : "simt_table_grad_long", bh * 2 * dd * RR * p->L,
                      qkv_bytes(bh, p->L, p->d, p->dtype, 1), s2, tc ? 1 : 2);
      if (tc) MLT_CUDA(simt_launch_table_grad_reduce(tl, s2));   // partials came from tc_bwd_q_kernel
      else MLT_CUDA(simt_launch_table_grad(tl, p->dtype, s2));
    }
    TableGradArgs tg{to_t4(p->global_q), p->G, wg.dallrel, wg.partial, wg.partial_bias, g->d_global_emb,
                     g->d_global_bias, p->B, p->H, p->R, p->d, simt_table_grad_chunks(p->G), p->scale};
    {
      ProfileScope ps(tc ? "table_grad_reduce_global" : "simt_table_grad_global", bh * 2 * dd * RR * p->G,
                      qkv_bytes(bh, p->G, p->d, p->dtype, 1), s2, tc ? 1 : 2);
      if (tc) MLT_CUDA(simt_launch_table_grad_reduce(tg, s2));
      else MLT_CUDA(simt_launch_table_grad(tg, p->dtype, s2));
    }
  }
  return MLT_OK;
}

// ---- long rows only (QkvRelativeLocalAttention) ---------------------------------------------
namespace {
int validate_local(const mlt_local_params* p) {
  if (!p) return MLT_ERR_NULL;
  if (p->abi_version != MLT_ABI_VERSION) return MLT_ERR_UNSUPPORTED;
  if (p->dtype != MLT_F32 && p->dtype != MLT_BF16) return MLT_ERR_DTYPE;
  if (p->B <= 0 || p->L <= 0 || p->G < 0 || p->H <= 0 || p->d <= 0 || p->R < 0 || p->local_radius < 1)
    return MLT_ERR_SHAPE;
  if (p->R > 64 || !simt_supports_head_dim(p->d)) return MLT_ERR_UNSUPPORTED;
  MLT_TRY(check_dropout(p->dropout_p));
  MLT_TRY(check_t4(p->q, p->dtype, p->d));
  MLT_TRY(check_t4(p->k, p->dtype, p->d));
  MLT_TRY(check_t4(p->v, p->dtype, p->d));
  MLT_TRY(check_t4(p->out, p->dtype, p->d));
  if (p->G > 0) {
    MLT_TRY(check_t4(p->side_k, p->dtype, p->d));
    MLT_TRY(check_t4(p->side_v, p->dtype, p->d));
  }
  if (!p->stats) return MLT_ERR_NULL;
  if ((p->tables.emb == nullptr) != (p->tables.bias == nullptr)) return MLT_ERR_NULL;
  if (p->R > 0 && !p->tables.emb) return MLT_ERR_NULL;
  if (p->side_mode == MLT_SIDE_COMPACT) {
    if (!p->example_ids || (p->G > 0 && !p->side_example_ids)) return MLT_ERR_NULL;
    if (p->R > 0 && p->G > 0 && !p->sentence_ids) return MLT_ERR_NULL;
    if (p->max_distance < 0) return MLT_ERR_SHAPE;
  } else if (p->side_mode != MLT_SIDE_EXPLICIT) {
    return MLT_ERR_UNSUPPORTED;
  }
  return MLT_OK;
}

// Reuse the global-local side builder through an equivalent mlt_gl_params view.
mlt_gl_params local_as_gl(const mlt_local_params* p) {
  mlt_gl_params g{};
  g.dtype = p->dtype; g.impl = p->impl;
  g.B = p->B; g.L = p->L; g.G = p->G > 0 ? p->G : 1; g.H = p->H; g.d = p->d; g.R = p->R;
  g.local_radius = p->local_radius;
  g.side_mode = p->side_mode;
  g.l2l_att_mask = p->att_mask; g.l2l_relative_att_ids = p->relative_att_ids;
  g.l2g_att_mask = p->side_att_mask; g.l2g_relative_att_ids = p->side_relative_att_ids;
  g.long_example_ids = p->example_ids; g.global_example_ids = p->side_example_ids;
  g.sentence_ids = p->sentence_ids; g.max_distance = p->max_distance;
  return g;
}

FwdArgs local_fwd_args(const mlt_local_params* p) {
  const mlt_gl_params g = local_as_gl(p);
  FwdArgs a{};
  a.rows = RowSet{to_t4(p->q), p->L, p->tables.emb, p->tables.bias, p->R};
  a.seg[0] = make_seg(p->k, p->v, p->L, 1, p->local_radius, gl_side(&g, L2L));
  a.nseg = 1;
  if (p->G > 0) {
    a.seg[1] = make_seg(p->side_k, p->side_v, p->G, 0, 0, gl_side(&g, L2G), p->L);
    a.nseg = 2;
  }
  a.out = to_t4(p->out);
  a.stats = p->stats;
  a.B = p->B; a.H = p->H; a.scale = p->scale; a.neg = p->neg;
  a.drop = make_dropout(p->dropout_p, p->dropout_seed, 0);
  return a;
}
}  // namespace

size_t mlt_local_workspace_bytes(const mlt_local_params* p, int bwd) {
  if (!p) return 0;
  size_t n = kAlign;
  if (bwd) n += row_ws_bytes(p->B, p->H, p->L, p->R, p->d) + tc_bwd_rows_ws_bytes(p->B, p->H, p->L, p->R);
  return n;
}

int mlt_local_rel_attn_fwd(const mlt_local_params* p, void* cuda_stream) {
  MLT_TRY(validate_local(p));
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const FwdArgs a = local_fwd_args(p);
  const bool tc = p->impl != MLT_IMPL_SIMT && tc_fwd_args_supported(a, p->dtype, p->d);
  if ((p->impl == MLT_IMPL_TC || p->impl == MLT_IMPL_TC_GENERIC) && !tc) return MLT_ERR_UNSUPPORTED;
  const double bh = (double)p->B * p->H;
  const double pairs = band_pairs(p->L, p->local_radius) + (double)p->L * p->G;
  return launch_fwd(a, tc, p->dtype, p->d, "fwd_local_rows", fwd_flops(bh, pairs, p->d, p->R, p->L),
                    qkv_bytes(bh, 4.0 * p->L + 2.0 * p->G, p->d, p->dtype, 1), st,
                    p->impl != MLT_IMPL_TC_GENERIC);
}

int mlt_local_rel_attn_bwd(const mlt_local_params* p, const mlt_local_grads* g, void* cuda_stream) {
  MLT_TRY(validate_local(p));
  if (!g) return MLT_ERR_NULL;
  MLT_TRY(check_t4(g->d_out, p->dtype, p->d));
  MLT_TRY(check_t4(g->d_q, p->dtype, p->d));
  MLT_TRY(check_t4(g->d_k, p->dtype, p->d));
  MLT_TRY(check_t4(g->d_v, p->dtype, p->d));
  if (p->G > 0) {
    MLT_TRY(check_t4(g->d_side_k, p->dtype, p->d));
    MLT_TRY(check_t4(g->d_side_v, p->dtype, p->d));
  }
  if (p->R > 0 && (!g->d_emb || !g->d_bias)) return MLT_ERR_NULL;
  if (!p->workspace || p->workspace_bytes < mlt_local_workspace_bytes(p, 1)) return MLT_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  char* wp = reinterpret_cast<char*>(align_up(reinterpret_cast<size_t>(p->workspace)));
  RowWs ws = carve_row_ws(wp, p->B, p->H, p->L, p->R, p->d);
  void* tcw[2] = {wp, wp};
  const FwdArgs f = local_fwd_args(p);
  BwdQArgs q{};
  q.rows = f.rows;
  q.seg[0] = f.seg[0];
  q.seg[1] = f.seg[1];
  q.nseg = f.nseg;
  q.out = to_t4(p->out); q.d_out = to_t4(g->d_out); q.d_q = to_t4(g->d_q);
  q.stats = p->stats; q.delta = ws.delta; q.allrel = ws.allrel; q.dallrel = ws.dallrel;
  q.B = p->B; q.H = p->H; q.scale = p->scale; q.neg = p->neg;
  q.drop = f.drop;
  BwdKVArgs kl{};
  kl.k = to_t4(p->k); kl.v = to_t4(p->v); kl.d_k = to_t4(g->d_k); kl.d_v = to_t4(g->d_v);
  kl.len = p->L;
  kl.src[0] = make_src(f.rows, g->d_out, p->stats, ws, 1, p->local_radius, f.seg[0].side, 0, f.drop);
  kl.nsrc = 1;
  kl.B = p->B; kl.H = p->H; kl.scale = p->scale; kl.neg = p->neg;
  BwdKVArgs ks{};
  if (p->G > 0) {
    ks.k = to_t4(p->side_k); ks.v = to_t4(p->side_v); ks.d_k = to_t4(g->d_side_k); ks.d_v = to_t4(g->d_side_v);
    ks.len = p->G;
    ks.src[0] = make_src(f.rows, g->d_out, p->stats, ws, 0, 0, f.seg[1].side, p->L, f.drop);
    ks.nsrc = 1;
    ks.B = p->B; ks.H = p->H; ks.scale = p->scale; ks.neg = p->neg;
  }
  const bool tc = p->impl != MLT_IMPL_SIMT && tc_bwd_q_supported(q, p->dtype, p->d) && bwd_kv_on_tc(kl) &&
                  (p->G == 0 || bwd_kv_on_tc(ks));
  if ((p->impl == MLT_IMPL_TC || p->impl == MLT_IMPL_TC_GENERIC) && !tc) return MLT_ERR_UNSUPPORTED;
  const double bh = (double)p->B * p->H, dd = p->d;
  const double p_l = band_pairs(p->L, p->local_radius), p_s = (double)p->L * p->G;
  MLT_TRY(launch_bwd_q(q, tc, tcw[0], p->dtype, p->d, "bwd_q_local_rows",
                       bh * (4 * dd * (p_l + p_s) + 2 * dd * p->R * p->L),
                       qkv_bytes(bh, 6.0 * p->L + 2.0 * p->G, p->d, p->dtype, 1), st));
  MLT_TRY(launch_bwd_kv(kl, tc, tcw, p->dtype, p->d, "bwd_kv_local_keys", bh * 4 * dd * p_l,
                        qkv_bytes(bh, 6.0 * p->L, p->d, p->dtype, 1), st));
  if (p->G > 0)
    MLT_TRY(launch_bwd_kv(ks, tc, tcw, p->dtype, p->d, "bwd_kv_side_keys", bh * 4 * dd * p_s,
                          qkv_bytes(bh, 2.0 * p->L + 4.0 * p->G, p->d, p->dtype, 1), st));
  if (p->R > 0) {
    TableGradArgs t{to_t4(p->q), p->L, ws.dallrel, ws.partial, ws.partial_bias, g->d_emb, g->d_bias,
                    p->B, p->H, p->R, p->d, simt_table_grad_chunks(p->L), p->scale};
    ProfileScope ps("simt_table_grad_local", bh * 2.0 * dd * p->R * p->L, qkv_bytes(bh, p->L, p->d, p->dtype, 1), st, 2);
    MLT_CUDA(simt_launch_table_grad(t, p->dtype, st));
  }
  return MLT_OK;
}

}  // extern "C"
