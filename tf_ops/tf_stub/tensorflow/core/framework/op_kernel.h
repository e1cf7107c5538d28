// Minimal stand-in for the TensorFlow op-kernel API -- TEST INFRASTRUCTURE, not TensorFlow.
//
// TensorFlow's headers are absent from this image, so tf_ops/davi_tf_ops.cc cannot be compiled against the real
// ones here.  This header tree declares exactly the part of the public API the shim uses (names, argument types
// and return types as documented for TF 2.x: OpKernel / OpKernelConstruction / OpKernelContext, Tensor,
// TensorShape, Status, errors::*, REGISTER_OP's builder, shape_inference::InferenceContext,
// REGISTER_KERNEL_BUILDER), so that `g++ -fsyntax-only -Itf_ops/tf_stub -Iinclude tf_ops/davi_tf_ops.cc`
// (tests/test_tf_shim.py) type-checks every C-ABI call site of the shim against include/davi.h: argument
// count, pointer constness, integer widths.  Nothing here is linked or executed.
#pragma once
#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <string>
#include <vector>

namespace tensorflow {

using int64 = long long;
using int32 = int;
enum DataType { DT_FLOAT = 1 };
constexpr const char* DEVICE_GPU = "GPU";

class Status {
 public:
  Status() = default;
  static Status OK() { return Status(); }
  bool ok() const { return ok_; }
  explicit Status(bool ok) : ok_(ok) {}
 private:
  bool ok_ = true;
};

namespace errors {
template <typename... Args> Status Internal(Args...) { return Status(false); }
template <typename... Args> Status InvalidArgument(Args...) { return Status(false); }
}  // namespace errors

class TensorShape {
 public:
  TensorShape() = default;
  TensorShape(std::initializer_list<int64> dims) : dims_(dims) {}
  int dims() const { return (int)dims_.size(); }
  int64 dim_size(int i) const { return dims_[i]; }
  void set_dim(int i, int64 v) { dims_[i] = v; }
  int64 num_elements() const { int64 n = 1; for (int64 d : dims_) n *= d; return n; }
 private:
  std::vector<int64> dims_;
};

template <typename T> struct FlatView {          // stands for Eigen::TensorMap: only .data() is used by the shim
  T* ptr;
  T* data() const { return ptr; }
};

class Tensor {
 public:
  int dims() const { return shape_.dims(); }
  int64 dim_size(int i) const { return shape_.dim_size(i); }
  const TensorShape& shape() const { return shape_; }
  int64 NumElements() const { return shape_.num_elements(); }
  template <typename T> FlatView<const T> flat() const { return FlatView<const T>{static_cast<const T*>(buf_)}; }
  template <typename T> FlatView<T> flat() { return FlatView<T>{static_cast<T*>(buf_)}; }
 private:
  TensorShape shape_;
  void* buf_ = nullptr;
};

struct GpuDeviceStub {                            // stands for Eigen::GpuDevice
  void* stream() const { return nullptr; }        // cudaStream_t in the real API
};

class OpKernelConstruction {
 public:
  template <typename T> Status GetAttr(const char* name, T* value) const { (void)name; (void)value; return Status::OK(); }
  void CtxFailure(const Status&) {}
  void CtxFailureWithWarning(const Status&) {}
};

class OpKernelContext {
 public:
  const Tensor& input(int index) const { (void)index; return dummy_; }
  Status allocate_output(int index, const TensorShape& shape, Tensor** out) { (void)index; (void)shape; *out = &dummy_; return Status::OK(); }
  Status allocate_temp(DataType type, const TensorShape& shape, Tensor* out) { (void)type; (void)shape; (void)out; return Status::OK(); }
  const GpuDeviceStub& eigen_gpu_device() const { return dev_; }
  void CtxFailure(const Status&) {}
  void CtxFailureWithWarning(const Status&) {}
 private:
  Tensor dummy_;
  GpuDeviceStub dev_;
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() = default;
  virtual void Compute(OpKernelContext* ctx) = 0;
};

// REGISTER_KERNEL_BUILDER(Name("Op").Device(DEVICE_GPU), KernelClass)
struct KernelDefBuilderStub {
  explicit KernelDefBuilderStub(const char*) {}
  KernelDefBuilderStub& Device(const char*) { return *this; }
};
inline KernelDefBuilderStub Name(const char* op) { return KernelDefBuilderStub(op); }
template <typename K> struct KernelRegistrarStub {
  explicit KernelRegistrarStub(const KernelDefBuilderStub&) { (void)sizeof(K); static_assert(sizeof(K) > 0, ""); }
  // instantiating the factory forces the kernel class to be complete and constructible from OpKernelConstruction*
  static OpKernel* Create(OpKernelConstruction* c) { return new K(c); }
};

}  // namespace tensorflow

#define DAVI_TFSTUB_CAT2(a, b) a##b
#define DAVI_TFSTUB_CAT(a, b) DAVI_TFSTUB_CAT2(a, b)
#define REGISTER_KERNEL_BUILDER(builder, ...)                                                         \
  static ::tensorflow::KernelRegistrarStub<__VA_ARGS__> DAVI_TFSTUB_CAT(davi_tfstub_kernel_, __COUNTER__)( \
      ::tensorflow::builder);                                                                          \
  static auto DAVI_TFSTUB_CAT(davi_tfstub_factory_, __COUNTER__) = &::tensorflow::KernelRegistrarStub<__VA_ARGS__>::Create

#define OP_REQUIRES(CTX, EXP, STATUS) \
  do {                                \
    if (!(EXP)) {                     \
      (CTX)->CtxFailure((STATUS));    \
      return;                         \
    }                                 \
  } while (0)
#define OP_REQUIRES_OK(CTX, ...)                       \
  do {                                                 \
    ::tensorflow::Status _s(__VA_ARGS__);              \
    if (!_s.ok()) {                                    \
      (CTX)->CtxFailureWithWarning(_s);                \
      return;                                          \
    }                                                  \
  } while (0)
#define TF_RETURN_IF_ERROR(...)                        \
  do {                                                 \
    ::tensorflow::Status _status = (__VA_ARGS__);      \
    if (!_status.ok()) return _status;                 \
  } while (0)
