// Device-side constructors of the explicit int32 side inputs.
// Replaces the per-example host tf.data construction of reference
// src/data/data_utils.py:305-332,335-379 and src/feature_utils.py:114-184: the O(S^2) int32
// tensors are produced where they are consumed instead of crossing PCIe.
// HBM-write-bound integer work: one thread per 4 consecutive columns, 128-bit stores.

#include "../../include/mlt_attn.h"
#include "mlt_common.cuh"

namespace {
using namespace mlt;

// out[b, i, c0..c0+3] for a row-major [B, rows, cols] tensor; cols % 4 handled by a scalar tail.
template <typename F>
__global__ void fill_rows_kernel(int32_t* out, int B, int rows, int cols, F f) {
  const int64_t quads_per_row = (cols + 3) / 4;
  const int64_t total = (int64_t)B * rows * quads_per_row;
  const bool vec_ok = (cols % 4 == 0);
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(t % quads_per_row) * 4;
    const int i = (int)((t / quads_per_row) % rows);
    const int b = (int)(t / (quads_per_row * rows));
    int32_t* dst = out + ((int64_t)b * rows + i) * cols + c0;
    if (vec_ok) {
      int4 v = make_int4(f(b, i, c0), f(b, i, c0 + 1), f(b, i, c0 + 2), f(b, i, c0 + 3));
      *reinterpret_cast<int4*>(dst) = v;
    } else {
      for (int c = 0; c < 4 && c0 + c < cols; ++c) dst[c] = f(b, i, c0 + c);
    }
  }
}

template <typename F>
cudaError_t fill_rows(int32_t* out, int B, int rows, int cols, F f, cudaStream_t st) {
  if (!out) return cudaSuccess;
  const int64_t total = (int64_t)B * rows * ((cols + 3) / 4);
  int blocks = (int)((total + 255) / 256);
  const int cap = 148 * 16;  // a few waves of the 148 SMs; grid-stride beyond that
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  fill_rows_kernel<<<blocks, 256, 0, st>>>(out, B, rows, cols, f);
  return cudaGetLastError();
}

// ---- recognition: explicit int32 side inputs -> compact descriptors (the inverse of the constructors) ----
// The Keras-level signature only carries the O(S^2) int32 tensors.  When they have the shape the
// reference's generators give them (masks = equality of segment labels, ids = the 1-D / 2-D / sentence
// rules), the compact descriptors reproduce them exactly and the attention kernels can rebuild masks and
// ids in registers.  Recognition derives candidate descriptors from O(S) slices and then compares EVERY
// element of every tensor with the rule's value: result[0] is 1 only when all of them are equal, so
// using the descriptors instead of the tensors cannot change a result.  HBM-read-bound integer work.

// res: [0] recognised, [1] max_distance, [2] mismatch flag (scratch), [3] reserved
__global__ void recog_init_kernel(int32_t* res, const int32_t* ids_at_minus_one, int given_distance) {
  res[0] = 0;
  res[1] = given_distance >= 0 ? given_distance : max(0, __ldg(ids_at_minus_one) - 1);   // id(-1) = D + 1
  res[2] = 0;
  res[3] = 0;
}
__global__ void recog_finish_kernel(int32_t* res) { res[0] = res[2] == 0 ? 1 : 0; }

// label[b, i] = index of the first non-zero entry of row (b, i) of mask [B, rows, cols], else -2 - offset - i
// (a label no other token carries).  `through` (optional) maps the index to the label of that column's token.
// hit[b, i] (optional) = first column whose ids entry equals res[1] * 2 + 2, else -1.  One warp per row.
// band (optional, [B, rows, 2 * radius + 1]): a row without any hit in `mask` takes the label of the first
// token of its block instead, found by walking the band towards earlier tokens (long tokens of an example
// that owns no global token, e.g. a padding tail: they still see one another).
__global__ void first_hit_rows_kernel(const int32_t* mask, const int32_t* ids, int B, int rows, int cols,
                                      const int32_t* through, int through_len, int unique_offset,
                                      int32_t* label, int32_t* hit, const int32_t* res,
                                      const int32_t* band = nullptr, int radius = 0) {
  const int lane = threadIdx.x & 31;
  const int64_t nrows = (int64_t)B * rows;
  const int want = ids ? 2 * __ldg(res + 1) + 2 : 0;
  for (int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); r < nrows;
       r += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int b = (int)(r / rows), i = (int)(r % rows);
    int first = -1, first_id = -1;
    for (int c0 = 0; c0 < cols && (first < 0 || (ids && first_id < 0)); c0 += 32) {
      const int c = c0 + lane;
      const bool m = c < cols && __ldg(mask + r * cols + c) != 0;
      const unsigned bm = __ballot_sync(0xffffffffu, m);
      if (first < 0 && bm) first = c0 + __ffs(bm) - 1;
      if (ids) {
        const bool e = c < cols && __ldg(ids + r * cols + c) == want;
        const unsigned be = __ballot_sync(0xffffffffu, e);
        if (first_id < 0 && be) first_id = c0 + __ffs(be) - 1;
      }
    }
    int peer = i;
    if (first < 0 && band) {
      // walk towards the start of the block: the first token this one sees, the first token THAT one sees, ...
      // (a block longer than the window needs more than one step; warp-uniform loop, bounded by the row length)
      const int bw = 2 * radius + 1;
      for (int step = 0; step < rows; ++step) {
        int nxt = peer;
        const int64_t pr = (int64_t)b * rows + peer;
        for (int c0 = 0; c0 < bw; c0 += 32) {
          const int c = c0 + lane;
          const bool m = c < bw && __ldg(band + pr * bw + c) != 0;
          const unsigned bm = __ballot_sync(0xffffffffu, m);
          if (bm) {
            nxt = peer + c0 + __ffs(bm) - 1 - radius;
            break;
          }
        }
        if (nxt >= peer || nxt < 0) break;   // sees nothing earlier (or a malformed mask): this is the block's first token
        peer = nxt;
      }
    }
    if (lane == 0) {
      int v = -2 - unique_offset - peer;
      if (first >= 0) v = through ? __ldg(through + (int64_t)b * through_len + first) : first;
      label[r] = v;
      if (hit) hit[r] = first_id;
    }
  }
}

// label[b, j] = index of the first row i with mask[b, i, j] != 0, else a unique label.  One thread per column.
__global__ void first_hit_cols_kernel(const int32_t* mask, int B, int rows, int cols, int unique_offset,
                                      int32_t* label) {
  const int64_t total = (int64_t)B * cols;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(t / cols), j = (int)(t % cols);
    const int32_t* m = mask + (int64_t)b * rows * cols + j;
    int first = -1;
    for (int i = 0; i < rows; ++i)
      if (__ldg(m + (int64_t)i * cols) != 0) { first = i; break; }
    label[t] = first >= 0 ? first : -2 - unique_offset - j;
  }
}

// in[b, i, c] == f(b, i, c) for every element, else *bad = 1.  Same traversal as fill_rows_kernel.
template <typename F>
__global__ void check_rows_kernel(const int32_t* in, int B, int rows, int cols, F f, int32_t* bad) {
  const int64_t quads_per_row = (cols + 3) / 4;
  const int64_t total = (int64_t)B * rows * quads_per_row;
  const bool vec_ok = (cols % 4 == 0) && (reinterpret_cast<uintptr_t>(in) % 16 == 0);
  bool ok = true;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(t % quads_per_row) * 4;
    const int i = (int)((t / quads_per_row) % rows);
    const int b = (int)(t / (quads_per_row * rows));
    const int32_t* src = in + ((int64_t)b * rows + i) * cols + c0;
    if (vec_ok) {
      const int4 v = __ldg(reinterpret_cast<const int4*>(src));
      ok &= v.x == f(b, i, c0) && v.y == f(b, i, c0 + 1) && v.z == f(b, i, c0 + 2) && v.w == f(b, i, c0 + 3);
    } else {
      for (int c = 0; c < 4 && c0 + c < cols; ++c) ok &= __ldg(src + c) == f(b, i, c0 + c);
    }
  }
  if (!ok) *bad = 1;
}

template <typename F>
cudaError_t check_rows(const int32_t* in, int B, int rows, int cols, F f, int32_t* bad, cudaStream_t st) {
  const int64_t total = (int64_t)B * rows * ((cols + 3) / 4);
  int blocks = (int)((total + 255) / 256);
  const int cap = 148 * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  check_rows_kernel<<<blocks, 256, 0, st>>>(in, B, rows, cols, f, bad);
  return cudaGetLastError();
}

int warp_rows_grid(int64_t nrows) {
  const int64_t blocks = (nrows + 7) / 8;   // 8 warps of a 256-thread block
  return (int)(blocks < 1 ? 1 : (blocks > 148 * 16 ? 148 * 16 : blocks));
}
}  // namespace

extern "C" {

int mlt_build_dense_side_inputs(const int32_t* example_ids, int32_t B, int32_t S,
                                mlt_id_layout layout, int32_t* att_mask,
                                int32_t* relative_att_ids, void* cuda_stream) {
  if (B <= 0 || S <= 0) return MLT_ERR_SHAPE;
  if (att_mask && !example_ids) return MLT_ERR_NULL;
  if (layout.max_distance < 0 || layout.num_patch_per_row < 0) return MLT_ERR_SHAPE;
  if (layout.num_patch_per_row > 0 &&
      (layout.num_core_layers <= 0 || layout.num_patch_per_row * layout.num_patch_per_row > S))
    return MLT_ERR_SHAPE;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  cudaError_t e = fill_rows(att_mask, B, S, S, [=] __device__(int b, int i, int j) -> int32_t {
    const int32_t* e_row = example_ids + (int64_t)b * S;
    return __ldg(e_row + i) == __ldg(e_row + j) ? 1 : 0;
  }, st);
  if (e != cudaSuccess) return (int)e;
  const int npr = layout.num_patch_per_row, core = layout.num_core_layers, D = layout.max_distance;
  e = fill_rows(relative_att_ids, B, S, S, [=] __device__(int, int i, int j) -> int32_t {
    return npr > 0 ? rel_id_2d(i, j, npr, core, D) : rel_id_1d(j - i, D);
  }, st);
  return e == cudaSuccess ? MLT_OK : (int)e;
}

int mlt_build_gl_side_inputs(const int32_t* long_example_ids, const int32_t* global_example_ids,
                             const int32_t* sentence_ids, int32_t B, int32_t L, int32_t G,
                             int32_t local_radius, int32_t max_distance, int32_t* const out[8],
                             void* cuda_stream) {
  if (!out) return MLT_ERR_NULL;
  if (B <= 0 || L <= 0 || G <= 0 || local_radius < 1 || max_distance < 0) return MLT_ERR_SHAPE;
  const bool need_e = out[0] || out[2] || out[4] || out[6];
  if (need_e && (!long_example_ids || !global_example_ids)) return MLT_ERR_NULL;
  if ((out[3] || out[7]) && !sentence_ids) return MLT_ERR_NULL;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int r = local_radius, D = max_distance, W = 2 * r + 1, voc = 2 * D + 1;
  const int32_t* le = long_example_ids;
  const int32_t* ge = global_example_ids;
  const int32_t* sid = sentence_ids;
  cudaError_t e;
#define MLT_FILL(ptr, rows, cols, body)                                                   \
  e = fill_rows(ptr, B, rows, cols, [=] __device__(int b, int i, int c) -> int32_t body, st); \
  if (e != cudaSuccess) return (int)e;
  MLT_FILL(out[0], L, W, {
    const int j = i + c - r;
    if (j < 0 || j >= L) return 0;
    return __ldg(le + (int64_t)b * L + i) == __ldg(le + (int64_t)b * L + j) ? 1 : 0;
  })
  MLT_FILL(out[1], L, W, { (void)b; (void)i; return rel_id_1d(c - r, D); })
  MLT_FILL(out[2], L, G, {
    return __ldg(le + (int64_t)b * L + i) == __ldg(ge + (int64_t)b * G + c) ? 1 : 0;
  })
  MLT_FILL(out[3], L, G, { return voc + (__ldg(sid + (int64_t)b * L + i) == c ? 1 : 0); })
  MLT_FILL(out[4], G, G, {
    return __ldg(ge + (int64_t)b * G + i) == __ldg(ge + (int64_t)b * G + c) ? 1 : 0;
  })
  MLT_FILL(out[5], G, G, { (void)b; return rel_id_1d(c - i, D); })
  MLT_FILL(out[6], G, L, {
    return __ldg(ge + (int64_t)b * G + i) == __ldg(le + (int64_t)b * L + c) ? 1 : 0;
  })
  MLT_FILL(out[7], G, L, { return voc + (__ldg(sid + (int64_t)b * L + c) == i ? 1 : 0); })
#undef MLT_FILL
  return MLT_OK;
}

int mlt_dense_compact_from_explicit(const int32_t* att_mask, const int32_t* relative_att_ids, int32_t B,
                                    int32_t S, mlt_id_layout hint, int32_t* q_example_ids,
                                    int32_t* k_example_ids, int32_t* result, void* cuda_stream) {
  if (!att_mask || !relative_att_ids || !q_example_ids || !k_example_ids || !result) return MLT_ERR_NULL;
  if (B <= 0 || S < 2 || hint.num_patch_per_row < 0) return MLT_ERR_SHAPE;
  if (hint.num_patch_per_row > 0 &&
      (hint.num_core_layers <= 0 || hint.max_distance < 0 || hint.num_patch_per_row * hint.num_patch_per_row > S))
    return MLT_ERR_SHAPE;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  // 1-D layout without a given distance: ids[0, 1, 0] is the id of offset -1
  recog_init_kernel<<<1, 1, 0, st>>>(result, relative_att_ids + S, hint.max_distance);
  first_hit_rows_kernel<<<warp_rows_grid((int64_t)B * S), 256, 0, st>>>(att_mask, nullptr, B, S, S, nullptr, 0, 0,
                                                                       q_example_ids, nullptr, result);
  {
    const int64_t total = (int64_t)B * S;
    first_hit_cols_kernel<<<(int)((total + 127) / 128), 128, 0, st>>>(att_mask, B, S, S, S, k_example_ids);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  int32_t* bad = result + 2;
  const int32_t* qe = q_example_ids;
  const int32_t* ke = k_example_ids;
  const int32_t* res = result;
  e = check_rows(att_mask, B, S, S, [=] __device__(int b, int i, int j) -> int32_t {
    return __ldg(qe + (int64_t)b * S + i) == __ldg(ke + (int64_t)b * S + j) ? 1 : 0;
  }, bad, st);
  if (e != cudaSuccess) return (int)e;
  const int npr = hint.num_patch_per_row, core = hint.num_core_layers;
  e = check_rows(relative_att_ids, B, S, S, [=] __device__(int, int i, int j) -> int32_t {
    const int D = __ldg(res + 1);
    return npr > 0 ? rel_id_2d(i, j, npr, core, D) : rel_id_1d(j - i, D);
  }, bad, st);
  if (e != cudaSuccess) return (int)e;
  recog_finish_kernel<<<1, 1, 0, st>>>(result);
  e = cudaGetLastError();
  return e == cudaSuccess ? MLT_OK : (int)e;
}

int mlt_gl_compact_from_explicit(const int32_t* const in[8], int32_t B, int32_t L, int32_t G,
                                 int32_t local_radius, int32_t* long_example_ids,
                                 int32_t* global_example_ids, int32_t* sentence_ids, int32_t* result,
                                 void* cuda_stream) {
  if (!in || !long_example_ids || !global_example_ids || !sentence_ids || !result) return MLT_ERR_NULL;
  for (int t = 0; t < 8; ++t)
    if (!in[t]) return MLT_ERR_NULL;
  if (B <= 0 || L <= 0 || G <= 0 || local_radius < 1) return MLT_ERR_SHAPE;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int r = local_radius, W = 2 * r + 1;
  // l2l ids do not depend on the row: entry r - 1 of row 0 is the id of offset -1
  recog_init_kernel<<<1, 1, 0, st>>>(result, in[1] + (r - 1), -1);
  // global labels: first global token each one may attend to; long labels: the label of the first global
  // token a long token may attend to; sentence: the global token whose l2g id is the "own sentence" id
  first_hit_rows_kernel<<<warp_rows_grid((int64_t)B * G), 256, 0, st>>>(in[4], nullptr, B, G, G, nullptr, 0, 0,
                                                                       global_example_ids, nullptr, result);
  first_hit_rows_kernel<<<warp_rows_grid((int64_t)B * L), 256, 0, st>>>(in[2], in[3], B, L, G, global_example_ids, G,
                                                                       G, long_example_ids, sentence_ids, result,
                                                                       in[0], r);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  int32_t* bad = result + 2;
  const int32_t* le = long_example_ids;
  const int32_t* ge = global_example_ids;
  const int32_t* sid = sentence_ids;
  const int32_t* res = result;
#define MLT_CHECK(ptr, rows, cols, body)                                                               \
  e = check_rows(ptr, B, rows, cols, [=] __device__(int b, int i, int c) -> int32_t body, bad, st);   \
  if (e != cudaSuccess) return (int)e;
  MLT_CHECK(in[0], L, W, {
    const int j = i + c - r;
    if (j < 0 || j >= L) return 0;
    return __ldg(le + (int64_t)b * L + i) == __ldg(le + (int64_t)b * L + j) ? 1 : 0;
  })
  MLT_CHECK(in[1], L, W, { (void)b; (void)i; return rel_id_1d(c - r, __ldg(res + 1)); })
  MLT_CHECK(in[2], L, G, {
    return __ldg(le + (int64_t)b * L + i) == __ldg(ge + (int64_t)b * G + c) ? 1 : 0;
  })
  MLT_CHECK(in[3], L, G, { return 2 * __ldg(res + 1) + 1 + (__ldg(sid + (int64_t)b * L + i) == c ? 1 : 0); })
  MLT_CHECK(in[4], G, G, {
    return __ldg(ge + (int64_t)b * G + i) == __ldg(ge + (int64_t)b * G + c) ? 1 : 0;
  })
  MLT_CHECK(in[5], G, G, { (void)b; return rel_id_1d(c - i, __ldg(res + 1)); })
  MLT_CHECK(in[6], G, L, {
    return __ldg(ge + (int64_t)b * G + i) == __ldg(le + (int64_t)b * L + c) ? 1 : 0;
  })
  MLT_CHECK(in[7], G, L, { return 2 * __ldg(res + 1) + 1 + (__ldg(sid + (int64_t)b * L + c) == i ? 1 : 0); })
#undef MLT_CHECK
  recog_finish_kernel<<<1, 1, 0, st>>>(result);
  e = cudaGetLastError();
  return e == cudaSuccess ? MLT_OK : (int)e;
}

}  // extern "C"
