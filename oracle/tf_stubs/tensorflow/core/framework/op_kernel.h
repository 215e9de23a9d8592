// Minimal stand-in for the four TensorFlow headers /root/reference/tf_ops/3d_interpolation/tf_interpolate.cpp includes
// (:6-9) -- TEST INFRASTRUCTURE (oracle/__init__.py).  Written from scratch for this repository: just enough of the op
// registration / OpKernel surface for that file to compile UNMODIFIED with g++ (oracle/build_ref.py), so that the
// reference's own three_nn / three_interpolate loops (and the shape checks of its OpKernels) can be executed here and
// on the GPU box without TensorFlow.  oracle/tf_stubs/harness.cpp drives the registered kernels through a C ABI.
#pragma once
#include <cstdint>
#include <functional>
#include <initializer_list>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace tensorflow {

struct Status {
  bool ok_ = true;
  std::string msg;
  static Status OK() { return Status(); }
  bool ok() const { return ok_; }
};

namespace errors {
inline Status InvalidArgument(const std::string& m) { Status s; s.ok_ = false; s.msg = m; return s; }
}  // namespace errors

struct TensorShape {
  std::vector<int64_t> d;
  TensorShape() {}
  TensorShape(std::initializer_list<int64_t> l) : d(l) {}
  int64_t dim_size(int i) const { return d[(size_t)i]; }
  int dims() const { return (int)d.size(); }
  int64_t num_elements() const { int64_t n = 1; for (auto v : d) n *= v; return n; }
};

template <class T>
struct Flat {
  T* p;
  T& operator()(int64_t i) const { return p[i]; }
};

// A tensor over caller-owned memory (inputs) or harness-owned memory (outputs).
struct Tensor {
  TensorShape shp;
  void* data = nullptr;
  int dims() const { return shp.dims(); }
  const TensorShape& shape() const { return shp; }
  template <class T> Flat<T> flat() { return Flat<T>{reinterpret_cast<T*>(data)}; }
  template <class T> Flat<const T> flat() const { return Flat<const T>{reinterpret_cast<const T*>(data)}; }
};

struct OpKernelConstruction {};

struct OpKernelContext {
  std::vector<Tensor> inputs;
  std::vector<Tensor> outputs;           // shapes filled by allocate_output; memory supplied by the harness
  std::vector<void*> out_buffers;        // caller-owned output memory, by output index
  Status status;
  const Tensor& input(int i) const { return inputs[(size_t)i]; }
  Status allocate_output(int i, const TensorShape& s, Tensor** t) {
    if ((size_t)i >= outputs.size()) outputs.resize((size_t)i + 1);
    outputs[(size_t)i].shp = s;
    outputs[(size_t)i].data = (size_t)i < out_buffers.size() ? out_buffers[(size_t)i] : nullptr;
    if (outputs[(size_t)i].data == nullptr) return errors::InvalidArgument("no buffer for output");
    *t = &outputs[(size_t)i];
    return Status::OK();
  }
  void SetStatus(const Status& s) { if (status.ok()) status = s; }
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext* context) = 0;
};

#define OP_REQUIRES(CTX, COND, STATUS) \
  do { if (!(COND)) { (CTX)->SetStatus(STATUS); return; } } while (0)
#define OP_REQUIRES_OK(CTX, EXPR) \
  do { ::tensorflow::Status s__ = (EXPR); if (!s__.ok()) { (CTX)->SetStatus(s__); return; } } while (0)

// ---- op registration (shape functions are accepted and ignored) ----
namespace shape_inference {
struct ShapeHandle {};
struct DimensionHandle {};
struct InferenceContext {
  ShapeHandle input(int) { return ShapeHandle(); }
  void set_output(int, ShapeHandle) {}
  Status WithRank(ShapeHandle, int, ShapeHandle*) { return Status::OK(); }
  DimensionHandle Dim(ShapeHandle, int) { return DimensionHandle(); }
  ShapeHandle MakeShape(std::initializer_list<DimensionHandle>) { return ShapeHandle(); }
};
}  // namespace shape_inference

struct OpDefBuilderStub {
  OpDefBuilderStub& Input(const char*) { return *this; }
  OpDefBuilderStub& Output(const char*) { return *this; }
  OpDefBuilderStub& Attr(const char*) { return *this; }
  OpDefBuilderStub& SetShapeFn(std::function<Status(shape_inference::InferenceContext*)>) { return *this; }
};
#define LRG_STUB_CAT2(a, b) a##b
#define LRG_STUB_CAT(a, b) LRG_STUB_CAT2(a, b)
#define REGISTER_OP(NAME) static ::tensorflow::OpDefBuilderStub LRG_STUB_CAT(lrg_stub_op_, __COUNTER__) = ::tensorflow::OpDefBuilderStub()

// ---- kernel registration: name -> factory ----
constexpr const char* DEVICE_CPU = "CPU";
constexpr const char* DEVICE_GPU = "GPU";
struct KernelName {
  std::string name;
  KernelName& Device(const char*) { return *this; }
  KernelName& HostMemory(const char*) { return *this; }
};
inline KernelName Name(const char* n) { return KernelName{n}; }
using KernelFactory = std::function<OpKernel*()>;
inline std::map<std::string, KernelFactory>& kernel_registry() {
  static std::map<std::string, KernelFactory> r;
  return r;
}
struct KernelRegistrar {
  KernelRegistrar(const KernelName& n, KernelFactory f) { kernel_registry()[n.name] = f; }
};
#define REGISTER_KERNEL_BUILDER(NAME, CLS)                                                          \
  static ::tensorflow::KernelRegistrar LRG_STUB_CAT(lrg_stub_kernel_, __COUNTER__)(                 \
      NAME, []() -> ::tensorflow::OpKernel* { static ::tensorflow::OpKernelConstruction c; return new CLS(&c); })

}  // namespace tensorflow
