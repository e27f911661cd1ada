// TensorFlow custom-op shim over the C ABI (include/mlt_attn.h).  SOURCE ONLY: this image has no
// TensorFlow headers, so the file is not part of the build; the tested binding is the torch one
// (ops.py).  Build recipe on a TF 2.5 machine:
//   g++ -std=c++14 -shared -fPIC mlt_gl_attn_op.cc -o _mlt_gl_attn_op.so \
//       $(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags()))') \
//       $(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_link_flags()))') \
//       -I../../include -L.. -lmlt_attn -DGOOGLE_CUDA=1
//
// It is a thin shim: it allocates outputs / workspace through the TF allocator, fills
// mlt_gl_params with raw device pointers and strides, and enqueues on TF's compute stream.
// Compact side inputs (long_example_ids, global_example_ids, sentence_ids) are the op's int32
// inputs; the explicit eight-tensor form maps onto MLT_SIDE_EXPLICIT in the same way.
#define EIGEN_USE_GPU
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/util/gpu_kernel_helper.h"

#include <cmath>

#include "mlt_attn.h"

namespace tf = tensorflow;

REGISTER_OP("MltGlAttn")
    .Input("long_q: T").Input("long_k: T").Input("long_v: T")
    .Input("global_q: T").Input("global_k: T").Input("global_v: T")
    .Input("long_emb: T").Input("long_bias: T").Input("global_emb: T").Input("global_bias: T")
    .Input("long_example_ids: int32").Input("global_example_ids: int32").Input("sentence_ids: int32")
    .Output("long_out: T").Output("global_out: T")
    .Output("long_stats: float").Output("global_stats: float")
    .Attr("T: {float, bfloat16}")
    .Attr("local_radius: int").Attr("max_distance: int")
    .SetShapeFn([](tf::shape_inference::InferenceContext* c) {
      c->set_output(0, c->input(0));
      c->set_output(1, c->input(3));
      return tf::Status::OK();
    });

namespace {

template <typename T>
mlt_tensor4 View(const tf::Tensor& t) {  // [B, len, H, d], dense
  const int64_t len = t.dim_size(1), h = t.dim_size(2), d = t.dim_size(3);
  return mlt_tensor4{const_cast<T*>(t.flat<T>().data()), len * h * d, h * d, d};
}

template <typename T>
class MltGlAttnOp : public tf::OpKernel {
 public:
  explicit MltGlAttnOp(tf::OpKernelConstruction* ctx) : tf::OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("local_radius", &local_radius_));
    OP_REQUIRES_OK(ctx, ctx->GetAttr("max_distance", &max_distance_));
  }

  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& lq = ctx->input(0);
    const tf::Tensor& gq = ctx->input(3);
    OP_REQUIRES(ctx, lq.dims() == 4 && gq.dims() == 4,
                tf::errors::InvalidArgument("q/k/v must be [B, len, H, d]"));
    const int B = lq.dim_size(0), L = lq.dim_size(1), H = lq.dim_size(2), d = lq.dim_size(3);
    const int G = gq.dim_size(1), R = ctx->input(6).dim_size(0);
    tf::Tensor *lo, *go, *ls, *gs;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, lq.shape(), &lo));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, gq.shape(), &go));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, tf::TensorShape({B, H, L, 2}), &ls));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, tf::TensorShape({B, H, G, 2}), &gs));

    mlt_gl_params p = {};
    p.abi_version = MLT_ABI_VERSION;
    p.dtype = std::is_same<T, float>::value ? MLT_F32 : MLT_BF16;
    p.impl = MLT_IMPL_AUTO;
    p.B = B; p.L = L; p.G = G; p.H = H; p.d = d; p.R = R;
    p.local_radius = local_radius_;
    p.scale = 1.0f / std::sqrt(static_cast<float>(d));
    p.neg = -1e9f;
    p.long_q = View<T>(ctx->input(0)); p.long_k = View<T>(ctx->input(1)); p.long_v = View<T>(ctx->input(2));
    p.global_q = View<T>(ctx->input(3)); p.global_k = View<T>(ctx->input(4)); p.global_v = View<T>(ctx->input(5));
    p.long_out = View<T>(*lo); p.global_out = View<T>(*go);
    p.long_stats = ls->flat<float>().data(); p.global_stats = gs->flat<float>().data();
    p.long_tables = {ctx->input(6).flat<T>().data(), ctx->input(7).flat<T>().data()};
    p.global_tables = {ctx->input(8).flat<T>().data(), ctx->input(9).flat<T>().data()};
    p.side_mode = MLT_SIDE_COMPACT;
    p.long_example_ids = ctx->input(10).flat<tf::int32>().data();
    p.global_example_ids = ctx->input(11).flat<tf::int32>().data();
    p.sentence_ids = ctx->input(12).flat<tf::int32>().data();
    p.max_distance = max_distance_;
    tf::Tensor ws;
    const size_t nbytes = mlt_gl_workspace_bytes(&p, /*bwd=*/0);
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_UINT8, tf::TensorShape({static_cast<int64_t>(nbytes)}), &ws));
    p.workspace = ws.flat<tf::uint8>().data();
    p.workspace_bytes = nbytes;
    const int rc = mlt_gl_attn_fwd(&p, ctx->eigen_gpu_device().stream());
    OP_REQUIRES(ctx, rc == MLT_OK, tf::errors::Internal("mlt_gl_attn_fwd: ", mlt_strerror(rc)));
  }

 private:
  int local_radius_, max_distance_;
};

}  // namespace

REGISTER_KERNEL_BUILDER(Name("MltGlAttn").Device(tf::DEVICE_GPU).TypeConstraint<float>("T"), MltGlAttnOp<float>);
REGISTER_KERNEL_BUILDER(Name("MltGlAttn").Device(tf::DEVICE_GPU).TypeConstraint<tf::bfloat16>("T"),
                        MltGlAttnOp<tf::bfloat16>);
// MltGlAttnGrad is built the same way around mlt_gl_attn_bwd (inputs: the forward's inputs, outputs
// and stats plus d_long_out / d_global_out; outputs: six q/k/v gradients and four fp32 table
// gradients; workspace = mlt_gl_workspace_bytes(&p, 1)); the Python side registers it with
// @tf.RegisterGradient("MltGlAttn") -- see INTEGRATION.md.
