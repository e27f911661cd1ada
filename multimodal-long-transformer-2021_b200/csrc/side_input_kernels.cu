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

}  // extern "C"
