// TensorFlow custom ops over the C ABI (include/mlt_attn.h): thin shims, no arithmetic.
//
//   MltDenseRelAttn / MltDenseRelAttnGrad     QkvRelativeAttention.call core with the reference's own side inputs:
//                                             att_mask, relative_att_ids int32 [B,S,S] exactly as built by
//                                             reference src/input_utils.py:35-44 and passed at
//                                             src/modeling/models/mmt_encoder.py:220-224
//   MltGlAttn / MltGlAttnGrad                 FusedGlobalLocalAttention core, the eight explicit int32 tensors
//                                             (l2l / l2g / g2g / g2l masks and relative ids)
//   MltGlAttnCompact / MltGlAttnCompactGrad   same core, masks / ids rebuilt in-kernel from example ids + sentence ids
//
// Each kernel allocates outputs / statistics / workspace through the TF allocator, fills the C struct with raw
// device pointers and dense strides and enqueues on TF's compute stream; errors come back as int codes and are
// turned into tf::errors.  Attention-probability dropout: attr `dropout_rate` + an int64 `seed` input; the forward
// op outputs nothing extra -- the gradient op is handed the same seed and regenerates the mask.
//
// This image has no TensorFlow: the file is type-checked against a minimal stub of the TF API
// (tests/tf_stub/, tests/test_tf_shim.py) so that every use of the C ABI in it is compiled; a real build is
//   g++ -std=c++14 -shared -fPIC mlt_ops.cc -o _mlt_ops.so $(python -c 'import tensorflow as tf;
//       print(" ".join(tf.sysconfig.get_compile_flags() + tf.sysconfig.get_link_flags()))')
//       -I../../include -L.. -lmlt_attn -DGOOGLE_CUDA=1
#define EIGEN_USE_GPU
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/util/gpu_kernel_helper.h"

#include <cmath>
#include <cstdint>
#include <type_traits>

#include "mlt_attn.h"

namespace tf = tensorflow;
using tf::shape_inference::InferenceContext;

namespace {

template <typename T>
mlt_tensor4 View(const tf::Tensor& t) {  // [B, len, H, d], dense
  const int64_t len = t.dim_size(1), h = t.dim_size(2), d = t.dim_size(3);
  return mlt_tensor4{const_cast<T*>(t.flat<T>().data()), len * h * d, h * d, d};
}
template <typename T>
int DtypeEnum() { return std::is_same<T, float>::value ? MLT_F32 : MLT_BF16; }
const int32_t* Ids(const tf::Tensor& t) { return t.NumElements() ? t.flat<tf::int32>().data() : nullptr; }
uint64_t Seed(const tf::Tensor& t) { return static_cast<uint64_t>(t.scalar<tf::int64>()()); }

tf::Status Workspace(tf::OpKernelContext* ctx, size_t nbytes, tf::Tensor* ws, void** ptr) {
  TF_RETURN_IF_ERROR(ctx->allocate_temp(tf::DT_UINT8, tf::TensorShape({static_cast<int64_t>(nbytes)}), ws));
  *ptr = ws->flat<tf::uint8>().data();
  return tf::Status::OK();
}
#define MLT_OP_CALL(ctx, expr, what)                                                                  \
  do {                                                                                                \
    const int rc_ = (expr);                                                                           \
    OP_REQUIRES(ctx, rc_ == MLT_OK, tf::errors::Internal(what, ": ", mlt_strerror(rc_)));             \
  } while (0)

// ------------------------------------------------------------------------------------------------
// Dense: inputs q k v emb bias att_mask relative_att_ids seed
template <typename T>
void FillDense(tf::OpKernelContext* ctx, float rate, mlt_dense_params* p) {
  const tf::Tensor& q = ctx->input(0);
  const tf::Tensor& k = ctx->input(1);
  *p = mlt_dense_params{};
  p->abi_version = MLT_ABI_VERSION;
  p->dtype = DtypeEnum<T>();
  p->impl = MLT_IMPL_AUTO;
  p->B = q.dim_size(0); p->Lq = q.dim_size(1); p->Lk = k.dim_size(1); p->H = q.dim_size(2); p->d = q.dim_size(3);
  p->R = ctx->input(3).dim_size(0);
  p->scale = 1.0f / std::sqrt(static_cast<float>(p->d));
  p->neg = -1e9f;
  p->dropout_p = rate;
  p->dropout_seed = Seed(ctx->input(7));
  p->q = View<T>(q); p->k = View<T>(k); p->v = View<T>(ctx->input(2));
  p->tables = {ctx->input(3).flat<T>().data(), ctx->input(4).flat<T>().data()};
  p->side_mode = MLT_SIDE_EXPLICIT;
  p->att_mask = Ids(ctx->input(5));
  p->relative_att_ids = Ids(ctx->input(6));
}

template <typename T>
class MltDenseRelAttnOp : public tf::OpKernel {
 public:
  explicit MltDenseRelAttnOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("dropout_rate", &rate_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& q = ctx->input(0);
    OP_REQUIRES(ctx, q.dims() == 4, tf::errors::InvalidArgument("q/k/v must be [B, len, H, d]"));
    mlt_dense_params p;
    FillDense<T>(ctx, rate_, &p);
    tf::Tensor *out, *stats, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, q.shape(), &out));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({p.B, p.H, p.Lq, 2}), &stats));
    p.out = View<T>(*out);
    p.stats = stats->flat<float>().data();
    p.workspace_bytes = mlt_dense_workspace_bytes(&p, 0);
    OP_REQUIRES_OK(ctx, Workspace(ctx, p.workspace_bytes, &ws, &p.workspace));
    MLT_OP_CALL(ctx, mlt_dense_rel_attn_fwd(&p, ctx->eigen_gpu_device().stream()), "mlt_dense_rel_attn_fwd");
  }
 private:
  float rate_;
};

// Grad: the forward's 8 inputs, then out, stats, d_out.  Outputs: d_q d_k d_v d_emb d_bias (tables in fp32).
template <typename T>
class MltDenseRelAttnGradOp : public tf::OpKernel {
 public:
  explicit MltDenseRelAttnGradOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("dropout_rate", &rate_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    mlt_dense_params p;
    FillDense<T>(ctx, rate_, &p);
    p.out = View<T>(ctx->input(8));
    p.stats = const_cast<float*>(ctx->input(9).flat<float>().data());
    tf::Tensor *dq, *dk, *dv, *de, *db, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, ctx->input(0).shape(), &dq));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, ctx->input(1).shape(), &dk));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, ctx->input(2).shape(), &dv));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, ctx->input(3).shape(), &de));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(4, ctx->input(4).shape(), &db));
    mlt_dense_grads g = {};
    g.d_out = View<T>(ctx->input(10));
    g.d_q = View<T>(*dq); g.d_k = View<T>(*dk); g.d_v = View<T>(*dv);
    g.d_emb = de->flat<float>().data();
    g.d_bias = db->flat<float>().data();
    p.workspace_bytes = mlt_dense_workspace_bytes(&p, 1);
    OP_REQUIRES_OK(ctx, Workspace(ctx, p.workspace_bytes, &ws, &p.workspace));
    MLT_OP_CALL(ctx, mlt_dense_rel_attn_bwd(&p, &g, ctx->eigen_gpu_device().stream()), "mlt_dense_rel_attn_bwd");
  }
 private:
  float rate_;
};

// ------------------------------------------------------------------------------------------------
// Global-local.  Common inputs 0..9: long_q long_k long_v global_q global_k global_v long_emb long_bias
// global_emb global_bias.  Explicit form: inputs 10..17 = l2l_att_mask l2l_relative_att_ids l2g_att_mask
// l2g_relative_att_ids g2g_att_mask g2g_relative_att_ids g2l_att_mask g2l_relative_att_ids, 18 = seed.
// Compact form: inputs 10..12 = long_example_ids global_example_ids sentence_ids, 13 = seed.
template <typename T, bool kCompact>
void FillGl(tf::OpKernelContext* ctx, int radius, int max_distance, float rate, mlt_gl_params* p) {
  const tf::Tensor& lq = ctx->input(0);
  const tf::Tensor& gq = ctx->input(3);
  *p = mlt_gl_params{};
  p->abi_version = MLT_ABI_VERSION;
  p->dtype = DtypeEnum<T>();
  p->impl = MLT_IMPL_AUTO;
  p->B = lq.dim_size(0); p->L = lq.dim_size(1); p->H = lq.dim_size(2); p->d = lq.dim_size(3);
  p->G = gq.dim_size(1);
  p->R = ctx->input(6).dim_size(0);
  p->local_radius = radius;
  p->scale = 1.0f / std::sqrt(static_cast<float>(p->d));
  p->neg = -1e9f;
  p->dropout_p = rate;
  p->long_q = View<T>(ctx->input(0)); p->long_k = View<T>(ctx->input(1)); p->long_v = View<T>(ctx->input(2));
  p->global_q = View<T>(ctx->input(3)); p->global_k = View<T>(ctx->input(4)); p->global_v = View<T>(ctx->input(5));
  p->long_tables = {ctx->input(6).flat<T>().data(), ctx->input(7).flat<T>().data()};
  p->global_tables = {ctx->input(8).flat<T>().data(), ctx->input(9).flat<T>().data()};
  if (kCompact) {
    p->side_mode = MLT_SIDE_COMPACT;
    p->long_example_ids = Ids(ctx->input(10));
    p->global_example_ids = Ids(ctx->input(11));
    p->sentence_ids = Ids(ctx->input(12));
    p->max_distance = max_distance;
    p->dropout_seed = Seed(ctx->input(13));
  } else {
    p->side_mode = MLT_SIDE_EXPLICIT;
    p->l2l_att_mask = Ids(ctx->input(10)); p->l2l_relative_att_ids = Ids(ctx->input(11));
    p->l2g_att_mask = Ids(ctx->input(12)); p->l2g_relative_att_ids = Ids(ctx->input(13));
    p->g2g_att_mask = Ids(ctx->input(14)); p->g2g_relative_att_ids = Ids(ctx->input(15));
    p->g2l_att_mask = Ids(ctx->input(16)); p->g2l_relative_att_ids = Ids(ctx->input(17));
    p->dropout_seed = Seed(ctx->input(18));
  }
}

template <typename T, bool kCompact>
class MltGlAttnOp : public tf::OpKernel {
 public:
  explicit MltGlAttnOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("local_radius", &radius_));
    OP_REQUIRES_OK(c, c->GetAttr("dropout_rate", &rate_));
    max_distance_ = 0;
    if (kCompact) OP_REQUIRES_OK(c, c->GetAttr("max_distance", &max_distance_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    OP_REQUIRES(ctx, ctx->input(0).dims() == 4 && ctx->input(3).dims() == 4,
                tf::errors::InvalidArgument("q/k/v must be [B, len, H, d]"));
    mlt_gl_params p;
    FillGl<T, kCompact>(ctx, radius_, max_distance_, rate_, &p);
    tf::Tensor *lo, *go, *ls, *gs, ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, ctx->input(0).shape(), &lo));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, ctx->input(3).shape(), &go));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, tf::TensorShape({p.B, p.H, p.L, 2}), &ls));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, tf::TensorShape({p.B, p.H, p.G, 2}), &gs));
    p.long_out = View<T>(*lo); p.global_out = View<T>(*go);
    p.long_stats = ls->flat<float>().data(); p.global_stats = gs->flat<float>().data();
    p.workspace_bytes = mlt_gl_workspace_bytes(&p, 0);
    OP_REQUIRES_OK(ctx, Workspace(ctx, p.workspace_bytes, &ws, &p.workspace));
    MLT_OP_CALL(ctx, mlt_gl_attn_fwd(&p, ctx->eigen_gpu_device().stream()), "mlt_gl_attn_fwd");
  }
 private:
  int radius_, max_distance_;
  float rate_;
};

// Grad inputs: the forward's inputs (19 explicit / 14 compact), then long_out global_out long_stats global_stats
// d_long_out d_global_out.  Outputs: six q/k/v gradients (T) and four table gradients (fp32).
template <typename T, bool kCompact>
class MltGlAttnGradOp : public tf::OpKernel {
 public:
  explicit MltGlAttnGradOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("local_radius", &radius_));
    OP_REQUIRES_OK(c, c->GetAttr("dropout_rate", &rate_));
    max_distance_ = 0;
    if (kCompact) OP_REQUIRES_OK(c, c->GetAttr("max_distance", &max_distance_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    constexpr int kFwdInputs = kCompact ? 14 : 19;
    mlt_gl_params p;
    FillGl<T, kCompact>(ctx, radius_, max_distance_, rate_, &p);
    p.long_out = View<T>(ctx->input(kFwdInputs + 0));
    p.global_out = View<T>(ctx->input(kFwdInputs + 1));
    p.long_stats = const_cast<float*>(ctx->input(kFwdInputs + 2).flat<float>().data());
    p.global_stats = const_cast<float*>(ctx->input(kFwdInputs + 3).flat<float>().data());
    mlt_gl_grads g = {};
    g.d_long_out = View<T>(ctx->input(kFwdInputs + 4));
    g.d_global_out = View<T>(ctx->input(kFwdInputs + 5));
    tf::Tensor* out[10];
    for (int i = 0; i < 10; ++i) OP_REQUIRES_OK(ctx, ctx->allocate_output(i, ctx->input(i).shape(), &out[i]));
    g.d_long_q = View<T>(*out[0]); g.d_long_k = View<T>(*out[1]); g.d_long_v = View<T>(*out[2]);
    g.d_global_q = View<T>(*out[3]); g.d_global_k = View<T>(*out[4]); g.d_global_v = View<T>(*out[5]);
    g.d_long_emb = out[6]->flat<float>().data(); g.d_long_bias = out[7]->flat<float>().data();
    g.d_global_emb = out[8]->flat<float>().data(); g.d_global_bias = out[9]->flat<float>().data();
    tf::Tensor ws;
    p.workspace_bytes = mlt_gl_workspace_bytes(&p, 1);
    OP_REQUIRES_OK(ctx, Workspace(ctx, p.workspace_bytes, &ws, &p.workspace));
    MLT_OP_CALL(ctx, mlt_gl_attn_bwd(&p, &g, ctx->eigen_gpu_device().stream()), "mlt_gl_attn_bwd");
  }
 private:
  int radius_, max_distance_;
  float rate_;
};

}  // namespace

// ---- op registrations ------------------------------------------------------------------------------
// Recognition of generator-shaped explicit side inputs, once per batch: the eight int32 tensors of
// make_global_local_transformer_side_inputs -> compact descriptors + result int32[4] (result[0] = 1 iff the
// descriptors reproduce every element of all eight; result[1] = max_distance).  Feed MltGlAttnCompact under
// tf.cond(result[0] == 1, ...) and MltGlAttn otherwise.
class MltGlSideInputsToCompactOp : public tf::OpKernel {
 public:
  explicit MltGlSideInputsToCompactOp(tf::OpKernelConstruction* c) : tf::OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("local_radius", &radius_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& l2l = ctx->input(0);
    const tf::Tensor& l2g = ctx->input(2);
    OP_REQUIRES(ctx, l2l.dims() == 3 && l2g.dims() == 3, tf::errors::InvalidArgument("side inputs must be [B, rows, cols]"));
    const int64_t B = l2l.dim_size(0), L = l2l.dim_size(1), G = l2g.dim_size(2);
    const int32_t* in[8];
    for (int i = 0; i < 8; ++i) in[i] = Ids(ctx->input(i));
    tf::Tensor *le, *ge, *sid, *res;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B, L}), &le));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, tf::TensorShape({B, G}), &ge));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, tf::TensorShape({B, L}), &sid));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, tf::TensorShape({4}), &res));
    MLT_OP_CALL(ctx,
                mlt_gl_compact_from_explicit(in, static_cast<int32_t>(B), static_cast<int32_t>(L), static_cast<int32_t>(G),
                                             radius_, le->flat<tf::int32>().data(), ge->flat<tf::int32>().data(),
                                             sid->flat<tf::int32>().data(), res->flat<tf::int32>().data(),
                                             ctx->eigen_gpu_device().stream()),
                "mlt_gl_compact_from_explicit");
  }
 private:
  int radius_;
};

#define MLT_QKV_TABLES                                                                                  \
  .Input("long_q: T").Input("long_k: T").Input("long_v: T")                                             \
  .Input("global_q: T").Input("global_k: T").Input("global_v: T")                                       \
  .Input("long_emb: T").Input("long_bias: T").Input("global_emb: T").Input("global_bias: T")
#define MLT_EXPLICIT_SIDE                                                                               \
  .Input("l2l_att_mask: int32").Input("l2l_relative_att_ids: int32").Input("l2g_att_mask: int32")       \
  .Input("l2g_relative_att_ids: int32").Input("g2g_att_mask: int32").Input("g2g_relative_att_ids: int32") \
  .Input("g2l_att_mask: int32").Input("g2l_relative_att_ids: int32")
#define MLT_COMPACT_SIDE                                                                                \
  .Input("long_example_ids: int32").Input("global_example_ids: int32").Input("sentence_ids: int32")
#define MLT_GL_FWD_OUT                                                                                  \
  .Output("long_out: T").Output("global_out: T").Output("long_stats: float").Output("global_stats: float")
#define MLT_GL_GRAD_TAIL                                                                                \
  .Input("long_out: T").Input("global_out: T").Input("long_stats: float").Input("global_stats: float")  \
  .Input("d_long_out: T").Input("d_global_out: T")                                                      \
  .Output("d_long_q: T").Output("d_long_k: T").Output("d_long_v: T")                                    \
  .Output("d_global_q: T").Output("d_global_k: T").Output("d_global_v: T")                              \
  .Output("d_long_emb: float").Output("d_long_bias: float").Output("d_global_emb: float")               \
  .Output("d_global_bias: float")
#define MLT_GL_ATTRS .Attr("T: {float, bfloat16}").Attr("local_radius: int").Attr("dropout_rate: float = 0.0")

static tf::Status GlFwdShape(InferenceContext* c) {
  c->set_output(0, c->input(0));
  c->set_output(1, c->input(3));
  return tf::Status::OK();
}
static tf::Status GradShape10(InferenceContext* c) {
  for (int i = 0; i < 10; ++i) c->set_output(i, c->input(i));
  return tf::Status::OK();
}

REGISTER_OP("MltDenseRelAttn")
    .Input("q: T").Input("k: T").Input("v: T").Input("emb: T").Input("bias: T")
    .Input("att_mask: int32").Input("relative_att_ids: int32").Input("seed: int64")
    .Output("out: T").Output("stats: float")
    .Attr("T: {float, bfloat16}").Attr("dropout_rate: float = 0.0")
    .SetShapeFn([](InferenceContext* c) { c->set_output(0, c->input(0)); return tf::Status::OK(); });
REGISTER_OP("MltDenseRelAttnGrad")
    .Input("q: T").Input("k: T").Input("v: T").Input("emb: T").Input("bias: T")
    .Input("att_mask: int32").Input("relative_att_ids: int32").Input("seed: int64")
    .Input("out: T").Input("stats: float").Input("d_out: T")
    .Output("d_q: T").Output("d_k: T").Output("d_v: T").Output("d_emb: float").Output("d_bias: float")
    .Attr("T: {float, bfloat16}").Attr("dropout_rate: float = 0.0")
    .SetShapeFn([](InferenceContext* c) { for (int i = 0; i < 5; ++i) c->set_output(i, c->input(i)); return tf::Status::OK(); });
REGISTER_OP("MltGlAttn") MLT_QKV_TABLES MLT_EXPLICIT_SIDE .Input("seed: int64") MLT_GL_FWD_OUT MLT_GL_ATTRS
    .SetShapeFn(GlFwdShape);
REGISTER_OP("MltGlAttnGrad") MLT_QKV_TABLES MLT_EXPLICIT_SIDE .Input("seed: int64") MLT_GL_GRAD_TAIL MLT_GL_ATTRS
    .SetShapeFn(GradShape10);
REGISTER_OP("MltGlAttnCompact") MLT_QKV_TABLES MLT_COMPACT_SIDE .Input("seed: int64") MLT_GL_FWD_OUT MLT_GL_ATTRS
    .Attr("max_distance: int").SetShapeFn(GlFwdShape);
REGISTER_OP("MltGlAttnCompactGrad") MLT_QKV_TABLES MLT_COMPACT_SIDE .Input("seed: int64") MLT_GL_GRAD_TAIL MLT_GL_ATTRS
    .Attr("max_distance: int").SetShapeFn(GradShape10);

REGISTER_OP("MltGlSideInputsToCompact") MLT_EXPLICIT_SIDE
    .Output("long_example_ids: int32").Output("global_example_ids: int32").Output("sentence_ids: int32")
    .Output("result: int32").Attr("local_radius: int");
REGISTER_KERNEL_BUILDER(Name("MltGlSideInputsToCompact").Device(tf::DEVICE_GPU), MltGlSideInputsToCompactOp);

#define MLT_REGISTER(NAME, ...)                                                                          \
  REGISTER_KERNEL_BUILDER(Name(NAME).Device(tf::DEVICE_GPU).TypeConstraint<float>("T").HostMemory("seed"), \
                          __VA_ARGS__<float>);                                                           \
  REGISTER_KERNEL_BUILDER(Name(NAME).Device(tf::DEVICE_GPU).TypeConstraint<tf::bfloat16>("T").HostMemory("seed"), \
                          __VA_ARGS__<tf::bfloat16>)
template <typename T> using GlExplicitOp = MltGlAttnOp<T, false>;
template <typename T> using GlExplicitGradOp = MltGlAttnGradOp<T, false>;
template <typename T> using GlCompactOp = MltGlAttnOp<T, true>;
template <typename T> using GlCompactGradOp = MltGlAttnGradOp<T, true>;
MLT_REGISTER("MltDenseRelAttn", MltDenseRelAttnOp);
MLT_REGISTER("MltDenseRelAttnGrad", MltDenseRelAttnGradOp);
MLT_REGISTER("MltGlAttn", GlExplicitOp);
MLT_REGISTER("MltGlAttnGrad", GlExplicitGradOp);
MLT_REGISTER("MltGlAttnCompact", GlCompactOp);
MLT_REGISTER("MltGlAttnCompactGrad", GlCompactGradOp);
