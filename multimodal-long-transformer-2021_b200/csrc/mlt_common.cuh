// Internal problem descriptors shared by the SIMT and tcgen05 kernels.
//
// One "attention problem" is a query set (rows) that attends, with ONE joint softmax, to up
// to two key segments.  A segment is either a band (same sequence, |j - i| <= radius) or a
// dense block.  This covers every block of the reference path:
//   dense (A)        : rows = S tokens,  segments = { dense(S) }
//   long rows of (B) : rows = L tokens,  segments = { band(L, r) [l2l], dense(G) [l2g] }
//   global rows of(B): rows = G tokens,  segments = { dense(G) [g2g], dense(L) [g2l] }
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mlt {

enum MaskRule : int { MR_NONE = 0, MR_EXPLICIT = 1, MR_EXAMPLE_ID = 2 };
enum IdRule : int {
  IDR_NONE = 0,
  IDR_EXPLICIT = 1,
  IDR_1D = 2,           // RelativePositionGenerator rule on o = j - i
  IDR_CROSS_QSENT = 3,  // l2g: 2D+1 + (sentence_ids[b, i] == j)   (query is the long token)
  IDR_CROSS_KSENT = 4,  // g2l: 2D+1 + (sentence_ids[b, j] == i)   (key is the long token)
  IDR_2D = 5            // MmtRelativePositionGenerator (image patches first, then text)
};

struct T4 {  // [B, len, H, d] view, element strides
  void* ptr;
  int64_t sb, sl, sh;
};

// Mask / relative-id provider for one (query set, key segment) block.
struct Side {
  int mask_rule;
  int id_rule;
  const int32_t* mask;  // explicit [B, Lq, W], W contiguous
  const int32_t* ids;
  int64_t sb, sq;       // explicit strides (elements)
  const int32_t* q_eid; // [B, Lq]
  const int32_t* k_eid; // [B, Lk]
  const int32_t* sent;  // [B, L_long] sentence id of the long token
  int q_len, k_len, sent_len;
  int max_distance;     // D
  int npr, core;        // 2-D rule: patches per row, core layers
};

inline bool side_is_explicit(const Side& s) { return s.mask_rule == MR_EXPLICIT || s.id_rule == IDR_EXPLICIT; }

struct KeySeg {
  T4 k, v;
  int len;
  int band;    // 1 = band segment (keys are the query sequence), 0 = dense
  int radius;
  Side side;
  int col_base;  // position of key 0 on the row set's concatenated key axis (dropout counter)
};

// ---- attention-probability dropout ----------------------------------------------------------
// keep(i, col) is a pure function of (seed, batch, head, row set, query row i, column col), where
// col indexes the row's concatenated key axis (segment 0 first).  Forward, both backward passes, the
// SIMT and the tcgen05 kernels and the numpy restatement in tests/ all evaluate the same function, so
// the mask is never stored.  The per-(batch, head, row set) salt comes from the 32-bit finaliser "lowbias32"
// (mix32: two multiplies, three xor-shifts); the per-element hash over the salted linear counter is the lighter
// mix_elem (it runs once per score element in three kernels); keep iff hash >= thr with
// thr = floor(p * 2^32): the rate is exact to 2^-32.
struct Dropout {
  uint32_t thr;        // 0: dropout off
  float inv_keep;      // 1 / (1 - p)
  uint32_t seed_lo, seed_hi;
  uint32_t rowset;     // 0: dense / long rows, 1: global rows
};

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
// per-element hash: two multiplies, one xor-shift (the counter is already salted by a full mix32; the
// comparison with the threshold reads the high bits, which the final multiply mixes best)
__host__ __device__ __forceinline__ uint32_t mix_elem(uint32_t x) {
  x *= 0x9e3779b1U;
  x ^= x >> 15;
  x *= 0x85ebca77U;
  return x;
}
// salt of one (batch * H + head, row set) unit
__host__ __device__ __forceinline__ uint32_t dropout_salt(const Dropout& d, uint32_t bh) {
  return mix32(d.seed_lo ^ mix32(d.seed_hi + 0x9e3779b9U * (2u * bh + d.rowset + 1u)));
}
// per-row base: salt + i * 0x10001 (rows 65537 counters apart; columns add their index)
__host__ __device__ __forceinline__ uint32_t dropout_row_base(uint32_t salt, int i) {
  return salt + (uint32_t)i * 0x00010001U;
}
__host__ __device__ __forceinline__ bool dropout_keep(uint32_t row_base, int col, uint32_t thr) {
  return mix_elem(row_base + (uint32_t)col) >= thr;
}

struct RowSet {  // query rows + the tables of their attention core
  T4 q;
  int len;
  const void* emb;   // [R, H, d] or null
  const void* bias;  // [R, H]
  int R;
};

struct FwdArgs {
  RowSet rows;
  KeySeg seg[2];
  int nseg;
  T4 out;
  float* stats;  // [B, H, Lq, 2]
  int B, H;
  float scale, neg;
  Dropout drop;
};

// Backward, query-centric pass: dq + per-row dallrel bins; also publishes delta / allrel.
struct BwdQArgs {
  RowSet rows;
  KeySeg seg[2];
  int nseg;
  T4 out, d_out, d_q;
  const float* stats;
  float* delta;    // ws [B, H, Lq]
  float* allrel;   // ws [B, H, Lq, R]
  float* dallrel;  // ws [B, H, Lq, R]
  // tcgen05 path only: per-tile table-gradient partials [(b * ntile + tile) * H + h][R][d] and
  // [...][R] (layout of TableGradArgs::partial / partial_bias); dallrel is then not written.
  float* tg_partial;
  float* tg_partial_bias;
  int B, H;
  float scale, neg;
  Dropout drop;
};

// One query source as seen from a key set (key-centric pass).
struct QuerySource {
  RowSet rows;
  T4 d_out;
  const float* stats;
  const float* delta;
  const float* allrel;
  int band, radius;
  Side side;  // row = query index, col = key index (band: j - i + r)
  int col_base;   // position of this key set on the source rows' concatenated key axis (dropout counter)
  Dropout drop;   // of the source's row set
};

struct BwdKVArgs {
  T4 k, v, d_k, d_v;
  int len;
  QuerySource src[2];
  int nsrc;
  int B, H;
  float scale, neg;
};

struct TableGradArgs {
  T4 q;
  int len;
  const float* dallrel;  // [B, H, Lq, R]
  float* partial;        // ws [B * nchunk, H, R, d]
  float* partial_bias;   // ws [B * nchunk, H, R]
  float* d_emb;          // [R, H, d]
  float* d_bias;         // [R, H]
  int B, H, R, d, nchunk;
  float scale;
};

// ---- integer rules (bit-exact with feature_utils / the oracle) ---------------------------

__host__ __device__ __forceinline__ int rel_id_1d(int offset, int max_distance) {
  // o >= 0 -> min(o, D);  o < 0 -> D + min(-o, D)   (pinned by reference
  // src/feature_utils_test.py:64-72,95-108)
  return offset >= 0 ? (offset < max_distance ? offset : max_distance)
                     : max_distance + (-offset < max_distance ? -offset : max_distance);
}

__host__ __device__ __forceinline__ int rel_id_2d(int i, int j, int npr, int core, int D) {
  // reference src/feature_utils.py:78-82,89-184
  const int n_img = npr * npr;
  const int image_part_id = n_img + 8 + 2 * D + 1;
  const bool qi = i < n_img, kj = j < n_img;
  if (qi && kj) {
    const int dy = j / npr - i / npr;
    const int dx = j % npr - i % npr;
    const int dia = 2 * core + 1;
    const int ay = dy < 0 ? -dy : dy, ax = dx < 0 ? -dx : dx;
    if (ay <= core && ax <= core) {
      int f = (dy * dia + dx) % (dia * dia);
      return f < 0 ? f + dia * dia : f;
    }
    const int vert = dy < -core ? 0 : (dy > core ? 2 : 1);
    const int horz = dx < -core ? 0 : (dx > core ? 2 : 1);
    // (vert, horz) -> clockwise direction index starting at 'top' (reference
    // src/feature_utils.py:186-255 dict order): T 0, TR 1, R 2, BR 3, B 4, BL 5, L 6, TL 7.
    const int code = vert * 3 + horz;  // 0 TL, 1 T, 2 TR, 3 L, 4 -, 5 R, 6 BL, 7 B, 8 BR
    int k;
    switch (code) {
      case 0: k = 7; break;
      case 1: k = 0; break;
      case 2: k = 1; break;
      case 3: k = 6; break;
      case 5: k = 2; break;
      case 6: k = 5; break;
      case 7: k = 4; break;
      default: k = 3; break;  // 8 (code 4 is the core, handled above)
    }
    return dia * dia + k;
  }
  if (qi) return image_part_id + 1;  // image row, text column -> text_part_id
  if (kj) return image_part_id;      // text row, image column -> image_part_id
  return rel_id_1d(j - i, D);
}

// q-side / k-side scalars are fetched by the caller (hoisted per row / staged per chunk).
__device__ __forceinline__ void side_eval(const Side& s, int b, int i, int j, int col, int q_e,
                                          int k_e, int q_sent, int k_sent, bool& ok, int& id) {
  if (s.mask_rule == MR_NONE) {
    ok = true;
  } else if (s.mask_rule == MR_EXPLICIT) {
    ok = __ldg(s.mask + (int64_t)b * s.sb + (int64_t)i * s.sq + col) != 0;
  } else {
    ok = (q_e == k_e);
  }
  switch (s.id_rule) {
    case IDR_EXPLICIT:
      id = __ldg(s.ids + (int64_t)b * s.sb + (int64_t)i * s.sq + col);
      break;
    case IDR_1D:
      id = rel_id_1d(j - i, s.max_distance);
      break;
    case IDR_CROSS_QSENT:
      id = 2 * s.max_distance + 1 + (q_sent == j ? 1 : 0);
      break;
    case IDR_CROSS_KSENT:
      id = 2 * s.max_distance + 1 + (k_sent == i ? 1 : 0);
      break;
    case IDR_2D:
      id = rel_id_2d(i, j, s.npr, s.core, s.max_distance);
      break;
    default:
      id = -1;
  }
}

__device__ __forceinline__ bool side_needs_q_eid(const Side& s) { return s.mask_rule == MR_EXAMPLE_ID; }
__device__ __forceinline__ bool side_needs_q_sent(const Side& s) { return s.id_rule == IDR_CROSS_QSENT; }
__device__ __forceinline__ bool side_needs_k_sent(const Side& s) { return s.id_rule == IDR_CROSS_KSENT; }

// ---- element access ---------------------------------------------------------------------

template <typename T>
__device__ __forceinline__ float to_f32(T x);
template <>
__device__ __forceinline__ float to_f32<float>(float x) { return x; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }

template <typename T>
__device__ __forceinline__ T from_f32(float x);
template <>
__device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// 4 consecutive elements -> fp32 (pointer must be aligned to 4 elements).
template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
  const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}

template <typename T>
__device__ __forceinline__ void store4(T* p, float4 v);
template <>
__device__ __forceinline__ void store4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 raw;
  raw.x = *reinterpret_cast<uint32_t*>(&a);
  raw.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = raw;
}

template <typename T>
__device__ __forceinline__ const T* row_ptr(const T4& t, int b, int l, int h) {
  return reinterpret_cast<const T*>(t.ptr) + (int64_t)b * t.sb + (int64_t)l * t.sl + (int64_t)h * t.sh;
}
template <typename T>
__device__ __forceinline__ T* row_ptr_mut(const T4& t, int b, int l, int h) {
  return reinterpret_cast<T*>(t.ptr) + (int64_t)b * t.sb + (int64_t)l * t.sl + (int64_t)h * t.sh;
}

// ---- launchers implemented in simt_kernels.cu ------------------------------------------
cudaError_t simt_launch_fwd(const FwdArgs& a, int dtype, int d, cudaStream_t st);
cudaError_t simt_launch_bwd_q(const BwdQArgs& a, int dtype, int d, cudaStream_t st);
cudaError_t simt_launch_bwd_kv(const BwdKVArgs& a, int dtype, int d, cudaStream_t st);
cudaError_t simt_launch_table_grad(const TableGradArgs& a, int dtype, cudaStream_t st);
// stage 2 only (fixed-order sum of the partials): used when the tcgen05 backward produced them
cudaError_t simt_launch_table_grad_reduce(const TableGradArgs& a, cudaStream_t st);
int simt_table_grad_chunks(int len);
bool simt_supports_head_dim(int d);

}  // namespace mlt
