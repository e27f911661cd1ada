// Minimal stand-in for the part of the TensorFlow C++ API that tf_shim/mlt_ops.cc uses, so that the shim can
// be TYPE-CHECKED (g++ -fsyntax-only) in an image without TensorFlow.  TEST INFRASTRUCTURE: declarations only,
// signatures modelled on TF 2.5 (tensorflow/core/framework/op_kernel.h, tensor.h, op.h, shape_inference.h).
#pragma once
#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <string>

namespace tensorflow {
typedef std::int32_t int32;
typedef std::int64_t int64;
typedef std::uint8_t uint8;
struct bfloat16 { std::uint16_t value; };
enum DataType { DT_UINT8 = 4, DT_FLOAT = 1 };
extern const char* const DEVICE_GPU;

class Status {
 public:
  static Status OK();
  bool ok() const;
};
namespace errors {
template <typename... A> Status Internal(A... a);
template <typename... A> Status InvalidArgument(A... a);
}  // namespace errors

class TensorShape {
 public:
  TensorShape();
  TensorShape(std::initializer_list<int64_t> dims);
};
template <typename T> struct Flat { T* data() const; };
template <typename T> struct Scalar { T operator()() const; };
class Tensor {
 public:
  Tensor();
  int dims() const;
  int64_t dim_size(int i) const;
  int64_t NumElements() const;
  const TensorShape& shape() const;
  template <typename T> Flat<T> flat();
  template <typename T> Flat<const T> flat() const;
  template <typename T> Scalar<T> scalar() const;
};

struct GpuStreamHolder { void* stream() const; };
class OpKernelConstruction {
 public:
  template <typename T> Status GetAttr(const char* name, T* value) const;
  void CtxFailureWithWarning(const char*, int, const Status&);
};
class OpKernelContext {
 public:
  const Tensor& input(int i);
  Status allocate_output(int i, const TensorShape& s, Tensor** out);
  Status allocate_temp(DataType t, const TensorShape& s, Tensor* out);
  const GpuStreamHolder& eigen_gpu_device() const;
  void CtxFailureWithWarning(const char*, int, const Status&);
};
class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*);
  virtual ~OpKernel();
  virtual void Compute(OpKernelContext* ctx) = 0;
};

namespace shape_inference {
struct ShapeHandle {};
class InferenceContext {
 public:
  ShapeHandle input(int i);
  void set_output(int i, ShapeHandle s);
};
}  // namespace shape_inference

struct OpDefBuilderWrapper {
  explicit OpDefBuilderWrapper(const char* name);
  OpDefBuilderWrapper& Input(const char*);
  OpDefBuilderWrapper& Output(const char*);
  OpDefBuilderWrapper& Attr(const char*);
  template <typename F> OpDefBuilderWrapper& SetShapeFn(F f);
};
struct KernelDefBuilder {
  explicit KernelDefBuilder(const char*);
  KernelDefBuilder& Device(const char*);
  template <typename T> KernelDefBuilder& TypeConstraint(const char*);
  KernelDefBuilder& HostMemory(const char*);
};
inline KernelDefBuilder Name(const char* n) { return KernelDefBuilder(n); }
struct KernelRegistrar {
  template <typename Op> static int Make(const KernelDefBuilder&);
};
}  // namespace tensorflow

#define TF_STUB_CAT2(a, b) a##b
#define TF_STUB_CAT(a, b) TF_STUB_CAT2(a, b)
#define REGISTER_OP(name) \
  static ::tensorflow::OpDefBuilderWrapper TF_STUB_CAT(tf_stub_op_, __COUNTER__) = ::tensorflow::OpDefBuilderWrapper(name)
#define REGISTER_KERNEL_BUILDER(builder, ...) \
  static int TF_STUB_CAT(tf_stub_kernel_, __COUNTER__) = ::tensorflow::KernelRegistrar::Make<__VA_ARGS__>(::tensorflow::builder)
#define OP_REQUIRES_OK(ctx, expr)                                  \
  do {                                                             \
    ::tensorflow::Status s_ = (expr);                              \
    if (!s_.ok()) { (ctx)->CtxFailureWithWarning(__FILE__, __LINE__, s_); return; } \
  } while (0)
#define OP_REQUIRES(ctx, cond, status)                             \
  do {                                                             \
    if (!(cond)) { (ctx)->CtxFailureWithWarning(__FILE__, __LINE__, (status)); return; } \
  } while (0)
#define TF_RETURN_IF_ERROR(expr)                                   \
  do {                                                             \
    ::tensorflow::Status s_ = (expr);                              \
    if (!s_.ok()) return s_;                                       \
  } while (0)
