/*
 * mlt_attn.h -- C ABI of the B200-native relative / global-local attention core.
 *
 * The reference has NO native boundary: its attention is reached
 * through the Keras layer call
 *     etc_layers.RelativeTransformerLayers(...)(inputs, att_mask, relative_att_ids, training)
 * at reference src/modeling/models/mmt_encoder.py:124-135 (ctor) / :220-224 (call),
 * which bottoms out in etcmodel's QkvRelativeAttention / QkvRelativeLocalAttention /
 * FusedGlobalLocalAttention [UPSTREAM-RECALLED, package not vendored].  This header is
 * therefore the boundary a TF custom op (OpKernel shim) or the
 * torch binding in multimodal-long-transformer-2021_b200/ops.py binds:
 *
 *   mlt_dense_rel_attn_fwd/bwd  <->  QkvRelativeAttention.call(queries, keys, values,
 *                                    att_mask, relative_att_ids)          
 *   mlt_gl_attn_fwd/bwd         <->  the attention core of FusedGlobalLocalAttention.call(
 *                                    long_input, global_input, l2l_att_mask, g2g_att_mask,
 *                                    l2g_att_mask, g2l_att_mask, l2l_relative_att_ids,
 *                                    g2g_relative_att_ids, l2g_relative_att_ids,
 *                                    g2l_relative_att_ids)          
 *   mlt_build_*                 <->  make_segmented_att_mask / RelativePositionGenerator /
 *                                    MmtRelativePositionGenerator (reference
 *                                    src/data/data_utils.py:305-332, src/feature_utils.py:29-255)
 *                                                                   
 *
 * Conventions
 *   - Plain C: raw device pointers, sizes and element strides.  No torch / TF types.
 *   - Ownership: the caller allocates EVERY buffer (outputs, statistics, workspace).  The
 *     library never allocates or frees device memory and keeps no pointer after return.
 *   - Asynchronous: work is enqueued on the caller's CUDA stream (a cudaStream_t passed as
 *     void*); the call returns immediately.
 *   - Errors: int return. 0 = ok; < 0 = MLT_ERR_* (bad argument / unsupported shape);
 *     > 0 = a cudaError_t from the launch.  Nothing throws, nothing exits.
 *   - Threading: re-entrant from any host thread and for any device.  Global state is limited to
 *     per-device one-time kernel attribute setup (mutex-guarded), a by-value cache of TMA tensor
 *     maps, and a small per-device pool of side streams taken with try_lock (a call never waits for
 *     another call: when the pool is busy it runs on the caller's stream alone).
 *   - q/k/v/out tensors are [B, len, H, d] with d contiguous (heads NOT transposed, as
 *     ProjectAttentionHeads produces them); strides are in elements.
 *   - Semantics:  s = (q.k + allrel[id]) * scale + neg * (1 - mask),
 *     allrel[p] = q.E[p] + bias[p], ids outside [0, R) contribute 0 (one-hot lookup),
 *     p = softmax over ALL key segments of the row jointly; with dropout_p > 0 (training) the
 *     probabilities are dropped AFTER the softmax and the kept ones scaled by 1 / (1 - dropout_p)
 *     (reference attention_probs_dropout_prob, src/configs/encoders.py:87-88).  The keep mask is a
 *     counter-based hash of (dropout_seed, batch, head, row set, query row, key column) -- see
 *     csrc/mlt_common.cuh (dropout_keep) and its numpy restatement tests/dropout_ref.py -- so the
 *     backward regenerates it from the seed; nothing is stored.
 */
#ifndef MLT_ATTN_H_
#define MLT_ATTN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLT_ABI_VERSION 1

#if defined(__GNUC__)
#define MLT_API __attribute__((visibility("default")))
#else
#define MLT_API
#endif

/* ---- error codes ------------------------------------------------------------------- */
enum {
  MLT_OK = 0,
  MLT_ERR_NULL = -1,        /* required pointer is NULL                                  */
  MLT_ERR_SHAPE = -2,       /* non-positive / inconsistent dimension                     */
  MLT_ERR_UNSUPPORTED = -3, /* valid request this build has no kernel for (e.g. d > 128) */
  MLT_ERR_STRIDE = -4,      /* stride / alignment requirement violated                   */
  MLT_ERR_WORKSPACE = -5,   /* workspace missing or too small                            */
  MLT_ERR_DTYPE = -6,       /* unknown dtype enum                                        */
  MLT_ERR_DROPOUT = -7      /* dropout_p outside [0, 1)                                  */
};

/* ---- enums ------------------------------------------------------------------------- */
enum { MLT_F32 = 0, MLT_BF16 = 1 };                 /* dtype of q/k/v/out/tables/grads   */
enum { MLT_SIDE_EXPLICIT = 0, MLT_SIDE_COMPACT = 1 }; /* how masks / ids are supplied    */
enum {                                               /* backend selection                */
  MLT_IMPL_AUTO = 0,   /* tcgen05 kernels when (bf16, d == 64, R <= 64), else SIMT        */
  MLT_IMPL_SIMT = 1,   /* force the fp32-accumulate CUDA-core kernels                    */
  MLT_IMPL_TC = 2,     /* force tcgen05 kernels (MLT_ERR_UNSUPPORTED if not applicable)   */
  MLT_IMPL_TC_GENERIC = 3 /* as MLT_IMPL_TC, but skip the kernels specialised for compact
                             global-local side inputs (cross-check of the two tcgen05 paths)  */
};

/* [B, len, H, d] view, d contiguous; strides in elements. */
typedef struct {
  void* ptr;
  int64_t stride_b;
  int64_t stride_l;
  int64_t stride_h;
} mlt_tensor4;

/* Relative tables of one attention core: emb [R, H, d], bias [R, H] (contiguous, same dtype
 * as q).  Either both NULL (no relative term) or both set. */
typedef struct {
  const void* emb;
  const void* bias;
} mlt_rel_tables;

/* Compact description of a 2-D (image) + 1-D (text) id layout
 * (MmtRelativePositionGenerator, reference src/feature_utils.py:29-255).  num_patch_per_row
 * == 0 selects the plain 1-D rule. */
typedef struct {
  int32_t num_patch_per_row;
  int32_t num_core_layers;
  int32_t max_distance;       /* text_relative_pos_max_distance / relative_pos_max_distance */
} mlt_id_layout;

/* ---- contract (A): dense relative attention ------------------------------------------ */
typedef struct {
  int32_t abi_version;        /* MLT_ABI_VERSION */
  int32_t dtype;              /* MLT_F32 / MLT_BF16 */
  int32_t impl;               /* MLT_IMPL_* */
  int32_t B, Lq, Lk, H, d, R; /* R = relative_vocab_size (0 if no tables) */
  float scale;                /* 1/sqrt(d) */
  float neg;                  /* additive mask constant, honoured literally (reference: -1e9) */
  float dropout_p;            /* attention-probability dropout rate in [0, 1); 0 = off */
  uint64_t dropout_seed;      /* keep mask = hash(seed, b, h, row, key); pass the same seed to _bwd */
  mlt_tensor4 q, k, v;        /* inputs */
  mlt_tensor4 out;            /* output [B, Lq, H, d] */
  float* stats;               /* output [B, H, Lq, 2] = (m, sum_j exp(s_j - m)) with m >= row max a
                               * softmax reference (the row max itself on most paths); required */
  mlt_rel_tables tables;
  int32_t side_mode;          /* MLT_SIDE_* */
  /* MLT_SIDE_EXPLICIT: int32 [B, Lq, Lk] each, contiguous; NULL mask = all ones, NULL ids = no
   * relative term. */
  const int32_t* att_mask;
  const int32_t* relative_att_ids;
  /* MLT_SIDE_COMPACT: mask[b,i,j] = (q_example_ids[b,i] == k_example_ids[b,j]); ids from
   * id_layout with query position i and key position j (self-attention layout). */
  const int32_t* q_example_ids; /* [B, Lq] */
  const int32_t* k_example_ids; /* [B, Lk] */
  mlt_id_layout id_layout;
  void* workspace;            /* mlt_dense_workspace_bytes() bytes, 256-B aligned */
  size_t workspace_bytes;
} mlt_dense_params;

typedef struct {
  mlt_tensor4 d_out;          /* input  [B, Lq, H, d] */
  mlt_tensor4 d_q, d_k, d_v;  /* outputs, same dtype as q */
  float* d_emb;               /* output fp32 [R, H, d] (NULL if no tables) */
  float* d_bias;              /* output fp32 [R, H] */
} mlt_dense_grads;

/* ---- contract (B): global-local attention -------------------------------------------- */
typedef struct {
  int32_t abi_version;
  int32_t dtype;
  int32_t impl;
  int32_t B, L, G, H, d, R;
  int32_t local_radius;
  float scale;
  float neg;
  float dropout_p;            /* as in mlt_dense_params */
  uint64_t dropout_seed;
  mlt_tensor4 long_q, long_k, long_v;       /* [B, L, H, d] */
  mlt_tensor4 global_q, global_k, global_v; /* [B, G, H, d] */
  mlt_tensor4 long_out, global_out;         /* outputs */
  float* long_stats;          /* output [B, H, L, 2] */
  float* global_stats;        /* output [B, H, G, 2] */
  mlt_rel_tables long_tables;   /* used by long rows (l2l + l2g) */
  mlt_rel_tables global_tables; /* used by global rows (g2g + g2l) */
  int32_t side_mode;
  /* MLT_SIDE_EXPLICIT (the Keras-level drop-in tensors), all int32, contiguous:
   * l2l [B, L, 2r+1] (column k <-> key j = i + k - r), l2g [B, L, G], g2g [B, G, G],
   * g2l [B, G, L].  NULL mask = all ones; NULL ids = no relative term for that block. */
  const int32_t* l2l_att_mask;
  const int32_t* l2l_relative_att_ids;
  const int32_t* l2g_att_mask;
  const int32_t* l2g_relative_att_ids;
  const int32_t* g2g_att_mask;
  const int32_t* g2g_relative_att_ids;
  const int32_t* g2l_att_mask;
  const int32_t* g2l_relative_att_ids;
  /* MLT_SIDE_COMPACT: everything above is rebuilt in registers from O(L+G) descriptors,
   * bit-exact with make_global_local_transformer_side_inputs_from_example_ids:
   *   masks: example-id equality (+ in-range for l2l);
   *   l2l / g2g ids: 1-D rule with max_distance D;
   *   l2g[b,i,g] = g2l[b,g,i] = 2D+1 + (sentence_ids[b,i] == g). */
  const int32_t* long_example_ids;   /* [B, L] */
  const int32_t* global_example_ids; /* [B, G] */
  const int32_t* sentence_ids;       /* [B, L] */
  int32_t max_distance;
  void* workspace;            /* mlt_gl_workspace_bytes() bytes, 256-B aligned */
  size_t workspace_bytes;
} mlt_gl_params;

typedef struct {
  mlt_tensor4 d_long_out, d_global_out;                 /* inputs */
  mlt_tensor4 d_long_q, d_long_k, d_long_v;             /* outputs */
  mlt_tensor4 d_global_q, d_global_k, d_global_v;       /* outputs */
  float* d_long_emb;    /* fp32 [R, H, d] */
  float* d_long_bias;   /* fp32 [R, H] */
  float* d_global_emb;
  float* d_global_bias;
} mlt_gl_grads;

/* ---- long rows only: QkvRelativeLocalAttention ------------------------------------------- */
typedef struct {
  int32_t abi_version;
  int32_t dtype;
  int32_t impl;
  int32_t B, L, G, H, d, R;   /* G = number of side keys (0: none) */
  int32_t local_radius;
  float scale;
  float neg;
  float dropout_p;            /* as in mlt_dense_params */
  uint64_t dropout_seed;
  mlt_tensor4 q, k, v;        /* [B, L, H, d] */
  mlt_tensor4 side_k, side_v; /* [B, G, H, d]; ignored when G == 0 */
  mlt_tensor4 out;            /* [B, L, H, d] */
  float* stats;               /* [B, H, L, 2] */
  mlt_rel_tables tables;
  int32_t side_mode;
  /* MLT_SIDE_EXPLICIT: att_mask / relative_att_ids [B, L, 2r+1]; side_* [B, L, G] */
  const int32_t* att_mask;
  const int32_t* relative_att_ids;
  const int32_t* side_att_mask;
  const int32_t* side_relative_att_ids;
  /* MLT_SIDE_COMPACT: as in mlt_gl_params (l2l and l2g blocks) */
  const int32_t* example_ids;       /* [B, L] */
  const int32_t* side_example_ids;  /* [B, G] */
  const int32_t* sentence_ids;      /* [B, L] */
  int32_t max_distance;
  void* workspace;
  size_t workspace_bytes;
} mlt_local_params;

typedef struct {
  mlt_tensor4 d_out;                     /* input */
  mlt_tensor4 d_q, d_k, d_v;             /* outputs */
  mlt_tensor4 d_side_k, d_side_v;        /* outputs (G > 0) */
  float* d_emb;                          /* fp32 [R, H, d] */
  float* d_bias;                         /* fp32 [R, H] */
} mlt_local_grads;

/* ---- entry points -------------------------------------------------------------------- */

/* Library / ABI identification. */
MLT_API int mlt_abi_version(void);
MLT_API const char* mlt_strerror(int code);
/* 1 if the tcgen05 path would be used for these parameters under MLT_IMPL_AUTO. */
MLT_API int mlt_gl_uses_tensor_cores(const mlt_gl_params* p);
MLT_API int mlt_dense_uses_tensor_cores(const mlt_dense_params* p);

/* Workspace sizes (bytes) for forward (bwd == 0) or backward (bwd != 0); pointer fields of *p
 * are ignored.  The forward needs a workspace too: with MLT_SIDE_COMPACT and a 2-D id layout
 * (num_patch_per_row > 0) the dense entry points materialise one [Lq, Lk] int32 id plane in it
 * (the ids do not depend on the batch element), which the kernels then read like explicit ids. */
MLT_API size_t mlt_dense_workspace_bytes(const mlt_dense_params* p, int bwd);
MLT_API size_t mlt_gl_workspace_bytes(const mlt_gl_params* p, int bwd);

/* Contract (A).  Replaces QkvRelativeAttention.call (reached from reference
 * src/modeling/models/mmt_encoder.py:220-224). */
MLT_API int mlt_dense_rel_attn_fwd(const mlt_dense_params* p, void* cuda_stream);
/* Gradient of the above (the reference relies on TF autodiff, src/tasks/pretraining.py:292-296).
 * Needs p->out and p->stats as written by the forward call. */
MLT_API int mlt_dense_rel_attn_bwd(const mlt_dense_params* p, const mlt_dense_grads* g, void* cuda_stream);

/* Contract (B).  Replaces the core of FusedGlobalLocalAttention.call [UPSTREAM-RECALLED]. */
MLT_API int mlt_gl_attn_fwd(const mlt_gl_params* p, void* cuda_stream);
MLT_API int mlt_gl_attn_bwd(const mlt_gl_params* p, const mlt_gl_grads* g, void* cuda_stream);

/* Long rows only.  Replaces QkvRelativeLocalAttention.call(queries, keys, values, att_mask,
 * relative_att_ids, side_keys, side_values, side_att_mask, side_relative_att_ids)
 * [UPSTREAM-RECALLED]. */
MLT_API size_t mlt_local_workspace_bytes(const mlt_local_params* p, int bwd);
MLT_API int mlt_local_rel_attn_fwd(const mlt_local_params* p, void* cuda_stream);
MLT_API int mlt_local_rel_attn_bwd(const mlt_local_params* p, const mlt_local_grads* g, void* cuda_stream);

/* Device-side side-input constructors (write the explicit int32 tensors the Keras signature
 * carries; bit-exact with the host constructors).
 * mlt_build_dense_side_inputs: example_ids [B,S] -> att_mask [B,S,S] (reference
 *   src/data/data_utils.py:320-322) and relative_att_ids [B,S,S] (1-D rule, or the 2-D
 *   image+text rule of reference src/feature_utils.py:114-184).  Either output may be NULL. */
MLT_API int mlt_build_dense_side_inputs(const int32_t* example_ids, int32_t B, int32_t S,
                                mlt_id_layout layout, int32_t* att_mask,
                                int32_t* relative_att_ids, void* cuda_stream);
/* mlt_build_gl_side_inputs: compact descriptors -> the eight [UPSTREAM-RECALLED] tensors of
 *   make_global_local_transformer_side_inputs.  out[] order: l2l_mask, l2l_ids, l2g_mask,
 *   l2g_ids, g2g_mask, g2g_ids, g2l_mask, g2l_ids; NULL entries are skipped. */
MLT_API int mlt_build_gl_side_inputs(const int32_t* long_example_ids, const int32_t* global_example_ids,
                             const int32_t* sentence_ids, int32_t B, int32_t L, int32_t G,
                             int32_t local_radius, int32_t max_distance, int32_t* const out[8],
                             void* cuda_stream);

/* Recognition of generator-shaped explicit side inputs: the inverse of the two constructors above.
 * The Keras-level signatures (reference src/modeling/models/mmt_encoder.py:220-224; FusedGlobalLocalAttention.call
 * [UPSTREAM-RECALLED]) only carry the O(S^2) int32 tensors.  These calls derive candidate compact descriptors
 * from them and then compare EVERY element of every tensor with the value the compact rule gives; they are
 * meant to be called once per batch (the tensors are shared by all layers), after which every layer's
 * forward and backward can run with MLT_SIDE_COMPACT.
 *   result: device int32[4].  result[0] = 1 iff every element of every tensor is reproduced exactly
 *           (only then may the descriptors replace the tensors); result[1] = max_distance; [2], [3] scratch.
 *   Outputs are written whether or not recognition succeeds.  Asynchronous like everything else: the caller
 *   reads result[] after synchronising the stream.
 * mlt_dense_compact_from_explicit: att_mask, relative_att_ids int32 [B,S,S] (square self-attention) ->
 *   q_example_ids, k_example_ids [B,S].  hint.num_patch_per_row > 0 checks the ids against that 2-D layout
 *   (hint.max_distance / num_core_layers as in mlt_id_layout); num_patch_per_row == 0 checks the 1-D rule with
 *   hint.max_distance, or with the distance read from the ids themselves when hint.max_distance < 0.
 * mlt_gl_compact_from_explicit: in[] in the order of mlt_build_gl_side_inputs' out[] (all eight required)
 *   -> long_example_ids [B,L], global_example_ids [B,G], sentence_ids [B,L] (-1 = no global token of its own). */
MLT_API int mlt_dense_compact_from_explicit(const int32_t* att_mask, const int32_t* relative_att_ids,
                                    int32_t B, int32_t S, mlt_id_layout hint, int32_t* q_example_ids,
                                    int32_t* k_example_ids, int32_t* result, void* cuda_stream);
MLT_API int mlt_gl_compact_from_explicit(const int32_t* const in[8], int32_t B, int32_t L, int32_t G,
                                 int32_t local_radius, int32_t* long_example_ids,
                                 int32_t* global_example_ids, int32_t* sentence_ids, int32_t* result,
                                 void* cuda_stream);

/* ---- profiling facility (bench.py; off by default) ------------------------------------ */
typedef struct {
  char name[48];   /* kernel role, e.g. "fwd_long_rows" */
  float ms;        /* CUDA-event duration on the caller's stream */
  double flops;    /* algorithmic FLOPs of this launch ("Work accounting") */
  double bytes;    /* algorithmic HBM bytes of this launch */
} mlt_kernel_time;
/* When on, every kernel launch is bracketed by events on the caller's stream. */
MLT_API int mlt_profile_enable(int on);
/* Synchronises the recorded events, copies up to max_entries records, clears the list. */
MLT_API int mlt_profile_read(mlt_kernel_time* out, int max_entries);
/* Number of kernels this library has launched in this process (bench.py "gpu_launches"). */
MLT_API long long mlt_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MLT_ATTN_H_ */
